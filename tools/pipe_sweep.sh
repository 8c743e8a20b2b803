#!/bin/bash
# slab geometry of the pinned staging ring (csrc/hostpipe.cu) against the pageable end-to-end times
# PIPE_CFGS="4,8,256 16,3,1024" tools/pipe_sweep.sh      (slab MiB, slabs, piece KiB)
for cfg in ${PIPE_CFGS:-16,3,1024 4,8,512 4,8,256 4,6,512 2,8,256 8,6,512}; do
  IFS=, read -r mb ns kb <<< "$cfg"
  echo "== SLAB_MB=$mb NSLAB=$ns PIECE_KB=$kb"
  LLACUDA_SLAB_MB=$mb LLACUDA_NSLAB=$ns LLACUDA_PIECE_KB=$kb LLACUDA_PIPE_STATS=1 timeout 300 python tools/pageable_probe.py 2>&1 | grep -v "^pinned  \|identical"
done
