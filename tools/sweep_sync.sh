#!/bin/bash
# tuning sweep on the GPU box: rebuild the kernels with each sync level, time early/late regimes for block sizes
cd "$(dirname "$0")/.."
for lvl in 0 1 2 3; do
  make -C roki-fd_b200/csrc -s -j16 clean >/dev/null 2>&1
  make -C roki-fd_b200/csrc -s -j16 EXTRA=-DRKFD_SYNC_LEVEL=$lvl >/dev/null 2>&1
  for blk in 128 256; do
    echo "sync_level=$lvl block=$blk"
    RKFD_FORCE_BLOCK=$blk python tools/exp_step_time_vs_contact.py 2>&1 | grep "base_z 0.45" | sed -n '1p;7p' | cut -c1-80
  done
done
