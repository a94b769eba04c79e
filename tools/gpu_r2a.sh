#!/bin/bash
# round-2 GPU pass A (1 GPU): parity suites, sharded-on-one-GPU, configs at size, bench, memcheck of smoke
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2a_gpus.txt
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_prove.py tests/test_golden.py -m gpu -x -q > gpurun_out/r2a_t1.log 2>&1
echo "t1 rc=$?"; tail -5 gpurun_out/r2a_t1.log
timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x > gpurun_out/r2a_t2.log 2>&1
echo "t2 rc=$?"; tail -15 gpurun_out/r2a_t2.log
timeout 900 python -m pytest tests/test_gpu_big_configs.py -m gpu -q > gpurun_out/r2a_t3.log 2>&1
echo "t3 rc=$?"; tail -8 gpurun_out/r2a_t3.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err; cut -c1-1500 gpurun_out/r2a_bench.json
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r2a_memcheck.log
timeout 900 python -m pytest tests/test_gpu_msm_large.py -m gpu -q -x > gpurun_out/r2a_t4.log 2>&1
echo "t4 rc=$?"; tail -8 gpurun_out/r2a_t4.log
timeout 600 python bench.py --workload msm --msm-max-log 24 > gpurun_out/r2a_msm.json 2> gpurun_out/r2a_msm.err
echo "msm rc=$?"; tail -3 gpurun_out/r2a_msm.err; cut -c1-1200 gpurun_out/r2a_msm.json
