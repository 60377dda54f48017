#!/usr/bin/env python
"""Train a GNN teacher and store its log-probabilities (out.npz).  Same command line as the
reference's train_teacher.py; the SAGE / GCN forward runs on libglnn_b200.so.

    python train_teacher.py --exp_setting tran --teacher GCN --dataset cora --device 0
"""
from glnn_b200.cli import main

if __name__ == "__main__":
    main("teacher")
