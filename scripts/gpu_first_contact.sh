#!/bin/bash
# First-contact GPU check: each kernel group under its own timeout so a hang cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 900 "$@" > gpurun_out/$name.log 2>&1; echo "exit $?"; tail -n 25 gpurun_out/$name.log; }
run t_gap      python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "operator_golden or gap_"
run t_anderson python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "anderson_kernels or residual_kernel"
run t_hid_fp32 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hidden_layer and fp32"
run t_hid_tc   python -m pytest tests/test_gpu_parity.py -q -m gpu -k "hidden_layer and tc_split"
run t_denoiser python -m pytest tests/test_gpu_parity.py -q -m gpu -k "denoiser_vs_oracle or f_two_calls"
run t_solver   python -m pytest tests/test_gpu_parity.py -q -m gpu -k "deq_andersonexp or forward_iteration or per_iterate"
