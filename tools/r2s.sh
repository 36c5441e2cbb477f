#!/usr/bin/env bash
# round 2, GPU call S (1 GPU): where a 1/8 shard's 20-step wavefront spends its time
set -u
RTB_PROFILE_DEPTHS=1 SHARD_PROBE_PROFILE=1 timeout 300 python tools/shard_probe.py 8 20 2>&1 | grep -E "depth|shard" | tail -12
timeout 300 python tools/shard_probe.py 8 64 2>&1 | tail -1
RTB_TAIL_PATHS=0 timeout 300 python tools/shard_probe.py 8 20 2>&1 | tail -1
RTB_TAIL_PATHS=262144 timeout 300 python tools/shard_probe.py 8 20 2>&1 | tail -1
RTB_TRAV_KERNEL=simple timeout 300 python tools/shard_probe.py 8 20 2>&1 | tail -1
RTB_SHADOW_KERNEL=simple timeout 300 python tools/shard_probe.py 8 20 2>&1 | tail -1
RTB_GRAPH=0 timeout 300 python tools/shard_probe.py 8 20 2>&1 | tail -1
RTB_PROFILE_DEPTHS=1 SHARD_PROBE_PROFILE=1 timeout 300 python tools/shard_probe.py 1 8 2>&1 | grep -E "depth|shard" | tail -12
