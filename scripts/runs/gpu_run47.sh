#!/bin/bash
# run 47: mid-tile re-alignment of the producers (experiment): f16 / tf32 TOM with BIXCOR_UMMA_TILE_SYNC_KB = 0 / 64 / 24
mkdir -p gpurun_out
for kb in 0 64 24; do
  echo "== BIXCOR_UMMA_TILE_SYNC_KB=$kb" >> gpurun_out/sync_kb_probe.log
  BIXCOR_UMMA_TILE_SYNC_KB=$kb timeout 100 python scripts/f16_probe.py f16 2>&1 | grep '"beta": 6.0' >> gpurun_out/sync_kb_probe.log
done
cut -c1-230 gpurun_out/sync_kb_probe.log
