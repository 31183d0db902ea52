#!/bin/bash
mkdir -p gpurun_out
( python __graft_entry__.py --smoke 2>&1 | tail -1
  python tools/smoke_probe.py 16 34
  AGCN_NO_FUSED_STATS=1 python tools/smoke_probe.py 16 34
  AGCN_MIX_SCORE_SIMT=1 python tools/smoke_probe.py 16 34 ) > gpurun_out/sp_probe.log 2>&1; cat gpurun_out/sp_probe.log
