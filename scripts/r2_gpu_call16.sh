set -x
cd "$GRAFT_REPO_ROOT"
for ex in peer none peer-nccl; do
echo "== exchange $ex"
TRAY_BENCH_STRONG_EXCHANGE=$ex timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 scripts/strong_only.py c4 2>&1 | grep -E "^\{|Error|error" | cut -c1-900
done
