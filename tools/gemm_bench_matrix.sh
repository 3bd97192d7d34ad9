# per-launch GEMM time = (t(repeat=21) - t(repeat=1)) / 20; NM_GEMM_DBG: 1 no MMAs, 2 no loads, 4 no stores
for dbg in ${DBGS:-0 4}; do
  for rep in 1 21; do NM_GEMM_DBG=$dbg timeout 120 python tools/gemm_bench.py 75648 256 256 0 $rep; done
  for rep in 1 21; do NM_GEMM_DBG=$dbg timeout 120 python tools/gemm_bench.py 256 256 75648 1 $rep; done
done
