#!/bin/bash
# Round 2, first GPU call: everything ungated on hardware, the PyTorch-CUDA denominator at B=256, the smplx probe,
# and ncu --set full captures of the kernels VERDICT r01 names (fused LBS form 3/5, chain GEMMs + glue_bwd in situ).
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_gpu1.sh'
mkdir -p gpurun_out
# 0. is smplx anywhere on the GPU box? (pins or un-pins oracle/smplh_lbs.py)
{ echo "== python -c import smplx"; python -c "import smplx; print(smplx.__file__)" 2>&1 | tail -1
  echo "== pip list | grep -i smpl/chumpy"; python -m pip list 2>/dev/null | grep -i -E "smpl|chumpy|body"
  echo "== /opt/wheelhouse"; ls /opt/wheelhouse 2>/dev/null | grep -i -E "smpl|chumpy"
  echo "== baseline/_ref"; ls baseline/_ref 2>&1 | head -3
  echo "== find / -name 'smplx*'"; find / -xdev \( -name 'smplx*' -o -name 'SMPLH*' -o -name 'smplh*' \) -not -path '*/proc/*' 2>/dev/null | head; } > gpurun_out/r02a_smplx_probe.txt 2>&1
# 1. the whole GPU suite with nothing gated
(HB_TEST_UNVERIFIED=1 timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/r02a_gpu_tests_ungated.txt
# 2. the reference algorithm as eager PyTorch on this GPU, at the benchmark batch (the 20x denominator)
(timeout 500 python bench.py --port-cuda 64,256 --cpu-steps 3 2> gpurun_out/r02a_port_cuda.err) > gpurun_out/r02a_port_cuda.json
# 3. fused LBS kernel: launch list + ncu --set full
bash tools/ncu_lbs_form.sh 3 5 lbs_fuseg_kernel r02a_fuseg35 > gpurun_out/r02a_fuseg35.log 2>&1
# 4. chain kernels as they run today: --set full on a split-K GroupNorm GEMM, its reverse twin and glue_bwd
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'umma_gemm3_kernel|glue_bwd_kernel' -s 60 -c 12 -o gpurun_out/r02a_chain_set_full -f \
  python tools/run_rollout_once.py 8 > gpurun_out/r02a_chain_set_full.log 2>&1
# 5. in-situ per-kernel times of a step (CUPTI, nothing serialised)
(timeout 200 python tools/profile_step.py 256 60 2>&1 | tail -40) > gpurun_out/r02a_profile_step.txt
(timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --lbs-skin 3 --lbs-blend 5 2>gpurun_out/r02a_bench_s3b5.err) > gpurun_out/r02a_bench_s3b5.json
tail -n 30 gpurun_out/r02a_gpu_tests_ungated.txt
cat gpurun_out/r02a_port_cuda.json
cat gpurun_out/r02a_smplx_probe.txt
tail -n 5 gpurun_out/r02a_fuseg35.log
