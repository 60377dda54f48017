set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29531 tools/exp_push_rate.py > gpurun_out/exp_push_rate_n4.jsonl 2> gpurun_out/exp_push_rate_n4.err
cat gpurun_out/exp_push_rate_n4.jsonl; tail -5 gpurun_out/exp_push_rate_n4.err
