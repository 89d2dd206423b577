#!/bin/bash
# experiment: batch parts on separate streams x factor-kernel occupancy limit (smem pad)
for parts in 1 2 4; do for pad in 0 16000 32000 48000; do
  echo "PARTS=$parts PAD=$pad: $(ALMPC_PARTS=$parts ALMPC_FAC_SMEM_PAD=$pad python tools/prof_one2.py 2>&1 | tail -1)"
done; done
