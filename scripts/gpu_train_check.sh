python -m pytest tests/test_gpu_training.py -x -q 2>&1 | tail -15
python scripts/bench_train.py --steps 3 --warmup 2 --batch 2 2>&1 | tail -2
