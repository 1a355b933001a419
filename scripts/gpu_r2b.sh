#!/bin/bash
# One gpurun call (r2, second session): GPU parity tests, head-layer / fused-pair timings with a sweep of the
# sliding kernel's CTA oversubscription, bench line.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2b.sh tag'
tag=${1:-u1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -n 8 gpurun_out/${tag}_pytest.log
for o in 4 3 5 6 8; do
  echo "== oversub $o" >> gpurun_out/${tag}_head.log
  B200SEG_SLIDE_OVERSUB=$o timeout 120 python scripts/head_layer.py >> gpurun_out/${tag}_head.log 2>&1
  B200SEG_SLIDE_OVERSUB=$o timeout 120 python scripts/head_layer.py 16 64 >> gpurun_out/${tag}_head.log 2>&1
  B200SEG_SLIDE_OVERSUB=$o timeout 120 python scripts/fuse_bench.py >> gpurun_out/${tag}_head.log 2>&1
done
cat gpurun_out/${tag}_head.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -n 3 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json
