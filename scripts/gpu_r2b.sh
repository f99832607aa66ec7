#!/bin/bash
# round 2, second pass (2 GPUs): work items in output order + in-place block exchange + shared host window.  Parity at N=1, sharded parity at N=2,
# bench at N=1 and N=2, PCIe zero-copy probe.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_full_size.py::test_config3_full_size_is_byte_identical_with_the_sdk_bake --durations=5 2>&1 | tail -15 | tee gpurun_out/r2b_pytest.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err
tail -c 1500 gpurun_out/r2b_bench_n1.json; tail -5 gpurun_out/r2b_bench_n1.err
OMM_B200_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err
tail -c 2500 gpurun_out/r2b_bench_n2.json; grep -v "^\[omm-b200 trace\]" gpurun_out/r2b_bench_n2.err | tail -5; grep "exchange" gpurun_out/r2b_bench_n2.err | tail -4
nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/probes/zc_probe.cu -o /tmp/zc_probe && /tmp/zc_probe | tee gpurun_out/r2b_zc_probe.txt
