#!/bin/bash
# Round 2, GPU call 13: full GPU suite after the LBS form pruning (incl. run() end-to-end with the stage files), where the library-L-BFGS run at B=256 leaves the finite
# numbers (tools/diag_nonfinite.py), and one ncu --set full capture of the batched prior GEMMs (forward GroupNorm and reverse).
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^  " | tail -25 | cut -c1-400) > gpurun_out/r02m_tests.txt
tail -4 gpurun_out/r02m_tests.txt
for prec in tensor exact; do
  (timeout 400 python tools/diag_nonfinite.py 256 $prec torch gpurun_out/r02m_diag_$prec 2>gpurun_out/r02m_diag_$prec.err) > gpurun_out/r02m_diag_$prec.jsonl
  tail -2 gpurun_out/r02m_diag_$prec.jsonl | cut -c1-1500
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:umma_gemm3_kernel -s 12 -c 6 -o gpurun_out/r02m_prior_gemm -f \
  python tools/run_rollout_once.py 59 > gpurun_out/r02m_prior_gemm.log 2>&1
ncu -i gpurun_out/r02m_prior_gemm.ncu-rep --page details 2>/dev/null | grep -E "umma_gemm3_kernel|^    Duration|L2 Cache Throughput|DRAM Throughput|Compute \(SM\) Throughput|Executed Ipc Active|L1/TEX Hit|L2 Hit|Mem Busy|Max Bandwidth" | cut -c1-160
