#!/bin/bash
# round-2 GPU call 23 (1 GPU): ncu --set full of the software-pipelined variant (RRI_SP_PIPE=2), to see why it loses
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
RRI_SP_PIPE=2 bash tools/ncu_sparse.sh full r02c
