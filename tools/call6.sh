set -x
mkdir -p gpurun_out
for mb in 59 235; do
PUSH_MB=$mb timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/exp_push_rate.py >> gpurun_out/exp_push_rate_n8.jsonl 2>> gpurun_out/exp_push_rate_n8.err
done
cat gpurun_out/exp_push_rate_n8.jsonl; tail -3 gpurun_out/exp_push_rate_n8.err
