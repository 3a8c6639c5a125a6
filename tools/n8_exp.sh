export TSG_BREAKDOWN_MODES=5
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_breakdown.py 2>&1 | grep " ms$" | head -6
echo SUB2
TSG_FORCE_SUB=2 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/dist_breakdown.py 2>&1 | grep " ms$" | head -6
