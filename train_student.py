#!/usr/bin/env python
"""Distil a stored teacher into an MLP student.  Same command line as the reference's
train_student.py; every training / evaluation step runs on libglnn_b200.so.

    python train_student.py --exp_setting tran --teacher SAGE --student MLP --dataset cora \
        --out_t_path outputs --device 0
"""
from glnn_b200.cli import main

if __name__ == "__main__":
    main("student")
