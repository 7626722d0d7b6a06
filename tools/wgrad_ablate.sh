#!/bin/bash
# Ablation of the TMA-fed weight-gradient kernel on three layer shapes (run on the GPU box).
export LAYERS='64->64@64,256->256@16,512->512@8'
for cfg in "" "SRVP_WGRAD_DBG=16" "SRVP_WGRAD_DBG=8" "SRVP_WGRAD_DBG=9" "SRVP_WGRAD_DBG=10" "SRVP_WGRAD_DBG=11" "SRVP_WGRAD_ISSUERS=2" "SRVP_WGRAD_TMA=0"; do
  echo "== ${cfg:-default}"
  env $cfg timeout 120 python tests/dev_wgrad_layers.py 2>&1 | grep -v Warning
done
