set -x
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider -k sharded_forward 2>&1 | tail -3
