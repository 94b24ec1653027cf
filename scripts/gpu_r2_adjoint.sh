#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py -q -m gpu -x -k "masked_adjoint or training_step_gradients or train_driver_equals" > gpurun_out/pytest_adjoint.log 2>&1; echo "pytest exit $?"; tail -n 25 gpurun_out/pytest_adjoint.log
timeout 600 python scripts/bench_train.py --denoiser SimpleCNN --steps 3 --warmup 1 > gpurun_out/train_cnn.log 2>&1; echo "train cnn exit $?"; tail -n 2 gpurun_out/train_cnn.log | cut -c 1-600
