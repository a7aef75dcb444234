#!/bin/bash
# run 45: per-tile producer barrier on long-K Grams (BIXCOR_UMMA_TILE_SYNC): parity subset + TOM timings with / without
mkdir -p gpurun_out
BIXCOR_UMMA_TILE_SYNC=1 timeout 300 python -m pytest tests -q -m gpu -x -k "tom or sparse_gram or ozaki" > gpurun_out/pytest_sync.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_sync.log
for m in 0 1; do
  echo "== BIXCOR_UMMA_TILE_SYNC=$m" >> gpurun_out/sync_probe.log
  BIXCOR_UMMA_TILE_SYNC=$m timeout 200 python scripts/f16_probe.py f16 tf32 2>&1 | grep '"beta": 6.0' >> gpurun_out/sync_probe.log
done
tail -2 gpurun_out/pytest_sync.log; cut -c1-230 gpurun_out/sync_probe.log
