#!/bin/bash
# tuning aid: fused-DCN stage costs (DV_DCN_DEBUG stubs) and stage depth; usage: gpurun -- bash tools/probe_dcn.sh TAG
TAG=${1:-dcn}
for v in "" "DV_DCN_DEBUG=1" "DV_DCN_DEBUG=2" "DV_DCN_DEBUG=3" "DV_DCN_STAGES=3" "DV_DCN_STAGES=5"; do
  echo "== $v"; env $v python tools/layer_profile.py lore 16 | grep -E "batch|ida_2.node_1|ida_2.proj_1|ida_0.node_1" 
done > gpurun_out/${TAG}_probe.txt 2>&1
cat gpurun_out/${TAG}_probe.txt
