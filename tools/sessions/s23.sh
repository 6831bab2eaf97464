#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
exec > gpurun_out/s23.log 2>&1
timeout 300 python tools/chain_probe.py 32

