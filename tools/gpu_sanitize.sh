#!/bin/bash
# NOTE (round 2): this GPU pool refuses compute-sanitizer ("closed on this pool ... runs under it have left GPUs needing a
# reset"), every run below returns at once with that message.  The bounds / uninitialised-read checks that DO run are
# tests/test_gpu_guard.py (canary bands around every tensor the engine allocates).  The job is kept for boxes that allow it.
#
# compute-sanitizer pass over the kernels (memcheck / racecheck / synccheck / initcheck on K1-K4 in
# the gpurun test job).  Run under gpurun on one GPU; logs under gpurun_out/san_*.log, one summary line per run in
# gpurun_out/sanitizer_summary.txt (copied to profiles/ by hand).
#   * PYTORCH_NO_CUDA_MEMORY_CACHING=1: every torch tensor is its own cudaMalloc, so an access one element past a tensor
#     is an access outside an allocation (with the caching allocator it would land inside a recycled block and go unseen)
#     and initcheck sees memory as uninitialised until OUR kernels (or torch fills) write it.
#   * every run has its own timeout: a sanitized persistent kernel is 10-100x slower, a hang must not cost the box.
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
export PYTORCH_NO_CUDA_MEMORY_CACHING=1
SUMMARY=gpurun_out/sanitizer_summary.txt
$SAN --version | head -2 > $SUMMARY
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader >> $SUMMARY
run() {
  name=$1; tool=$2; limit=$3; shift 3
  log=gpurun_out/san_${name}_${tool}.log
  t0=$(date +%s)
  timeout $limit $SAN --tool $tool --print-limit 30 --error-exitcode 86 "$@" > $log 2>&1
  rc=$?
  t1=$(date +%s)
  tests=$(grep -E "^[0-9]+ (passed|failed)|passed|failed" $log | tail -1)
  verdict=$(grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $log | tail -1)
  echo "$name | $tool | rc=$rc | $((t1 - t0)) s | pytest: ${tests:-none} | ${verdict:-no summary line}" | tee -a $SUMMARY
}
PT="python -m pytest -q -m gpu -p no:cacheprovider"
K=tests/test_gpu_kernels.py
# K1 CSR build / graph offsets / collate / wire expansion, K2 aggregation, K3 GEMM, K4 pool + head, encoder front
for tool in memcheck racecheck synccheck initcheck; do
  run k1_csr $tool 300 $PT $K -k "csr or graph_ptr or collate or wire"
  run k2_aggregate $tool 400 $PT $K -k "aggregate"
  run k3_gemm $tool 400 $PT $K -k "gemm"
  run k4_pool_encoder $tool 300 $PT $K -k "pool or encoder"
done
# whole forwards (all kernels in sequence, ragged tails, pool-fused last layer) and one training step: memcheck only
run fwd memcheck 500 $PT tests/test_gpu_forward.py -k "graphsage_mean_6x512 or ragged or single_graph or isolated"
run poolfuse memcheck 400 $PT tests/test_gpu_poolfuse.py -k "block_flags or block_sums or random_ragged or batch_none"
run train memcheck 500 $PT tests/test_gpu_train.py -k "bn_batch_stats or sage_backward_rows or weight_gradient or sgemm_colsum or deterministic or max_aggregation"
cat $SUMMARY
