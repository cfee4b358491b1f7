#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
timeout 600 python -m pytest tests -q -m gpu -p no:cacheprovider -x -k "split_kv" 2>&1 | tail -8
timeout 600 python scripts/splitkv_timing.py 2>&1 | tail -12
