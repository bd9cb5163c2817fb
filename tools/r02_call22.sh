#!/bin/bash
# round-2 GPU call 22 (1 GPU): ncu --set full of the default sparse pass kernel (sp_pass_stream_kernel)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
bash tools/ncu_sparse.sh full r02b
