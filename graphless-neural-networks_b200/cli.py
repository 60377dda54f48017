"""Command-line drivers with the reference's flags, defaults, output tree and printed score line
(/root/reference/train_teacher.py:21-149,152-342 and train_student.py:22-165,168-384), so that
experiments/*.sh keep working unchanged.  The two scripts share one argument table and one run
skeleton; they contain no math -- everything numerical goes through train_and_eval / models."""
import argparse
from pathlib import Path

import numpy as np
import torch
import torch.optim as optim

from . import train_and_eval as TE
from .dataloader import load_data, load_out_t
from .models import Model
from .utils import (check_readable, check_writable, compute_min_cut_loss, feature_prop,
                    get_evaluator, get_logger, get_training_config, graph_split, set_seed)

# (flag, type, default, help) -- store_true flags have type None.  Order and defaults follow the
# reference; the student adds --student / --lamb / --out_t_path and a smaller --hidden_dim.
_COMMON = [
    ("--device", int, -1, "CUDA device, -1 means CPU"),
    ("--seed", int, 0, "Random seed"),
    ("--log_level", int, 20, "Logger levels for run {10: DEBUG, 20: INFO, 30: WARNING}"),
    ("--console_log", None, False, "Set to True to display log info in console"),
    ("--output_path", str, "outputs", "Path to save outputs"),
    ("--num_exp", int, 1, "Repeat how many experiments"),
    ("--exp_setting", str, "tran", "Experiment setting, one of [tran, ind]"),
    ("--eval_interval", int, 1, "Evaluate once per how many epochs"),
    ("--save_results", None, False, "Save the loss curves, trained model, and min-cut loss"),
    ("--dataset", str, "cora", "Dataset"),
    ("--data_path", str, "./data", "Path to data"),
    ("--labelrate_train", int, 20, "How many labeled data per class as train set"),
    ("--labelrate_val", int, 30, "How many labeled data per class in valid set"),
    ("--split_idx", int, 0, "For Non-Homo datasets only, one of [0,1,2,3,4]"),
    ("--model_config_path", str, "./train.conf.yaml", "Path to model configeration"),
    ("--teacher", str, "SAGE", "Teacher model"),
    ("--num_layers", int, 2, "Model number of layers"),
    ("--hidden_dim", int, 128, "Model hidden layer dimensions"),
    ("--dropout_ratio", float, 0, None),
    ("--norm_type", str, "none", "One of [none, batch, layer]"),
    ("--batch_size", int, 512, None),
    ("--fan_out", str, "5,5", "Number of samples for each layer in SAGE. Length = num_layers"),
    ("--num_workers", int, 0, "Number of workers for sampler"),
    ("--learning_rate", float, 0.01, None),
    ("--weight_decay", float, 0.0005, None),
    ("--max_epoch", int, 500, "Maximum number of epochs"),
    ("--patience", int, 50, "Early stop after this many evaluations without improvement"),
    ("--feature_noise", float, 0, "White-noise level added to the features, in [0, 1]"),
    ("--split_rate", float, 0.2, "Rate for graph split, see graph_split"),
    ("--compute_min_cut", None, False, "Compute and store the min-cut loss"),
    ("--feature_aug_k", int, 0, "Augment node features with feature_aug_k-hop propagated features"),
]
_STUDENT_EXTRA = [
    ("--student", str, "MLP", "Student model"),
    ("--lamb", float, 0, "Balances loss from hard labels and teacher outputs, in [0, 1]"),
    ("--out_t_path", str, "outputs", "Path to load teacher outputs"),
]


def build_parser(role):
    parser = argparse.ArgumentParser(description=f"GLNN {role} on the B200 kernels")
    table = list(_COMMON)
    if role == "student":
        table = [(f, t, 64 if f == "--hidden_dim" else d, h) for f, t, d, h in table] + _STUDENT_EXTRA
    for flag, typ, default, helptext in table:
        if typ is None:
            parser.add_argument(flag, action="store_true", help=helptext)
        else:
            parser.add_argument(flag, type=typ, default=default, help=helptext)
    return parser


def _result_dir(args, base, leaf):
    if args.exp_setting == "tran":
        return Path.cwd().joinpath(base, "transductive", args.dataset, leaf, f"seed_{args.seed}")
    if args.exp_setting == "ind":
        return Path.cwd().joinpath(base, "inductive", f"split_rate_{args.split_rate}", args.dataset,
                                   leaf, f"seed_{args.seed}")
    raise ValueError(f"Unknown experiment setting! {args.exp_setting}")


# Teacher log-probabilities of teacher runs made in THIS process, keyed by their output directory
# (SURVEY.md 8f row 3): a student run that follows in the same process (distill_pipeline, or two
# main() calls) takes the device tensor from here instead of reading out.npz back from disk.  The
# file is still written -- it is the reference's hand-off format (train_teacher.py:296-297).
_OUT_T_ON_DEVICE = {}


def _resolve_device(index):
    """The reference's --device default is -1 = CPU (train_teacher.py:26).  This implementation has no
    CPU path (every hot function is a CUDA kernel), so -1 means "the current CUDA device" here, and a
    machine without CUDA is refused up front -- before any output directory is created."""
    if not torch.cuda.is_available():
        raise RuntimeError("glnn_b200 needs a CUDA device (B200, sm_100a): there is no CPU path. "
                           "Run the reference itself for a CPU run.")
    return torch.device(f"cuda:{index}") if index >= 0 else torch.device("cuda", torch.cuda.current_device())


def run(args, role):
    """One seed.  Returns [score_test] (tran) or [score_test_tran, score_test_ind] (ind)."""
    student = role == "student"
    set_seed(args.seed)
    device = _resolve_device(args.device)

    # the reference rewrites output_path (and the model name) for the ablations only when seed == 0,
    # so that repeat_run's later seeds inherit the rewritten values
    if args.feature_noise != 0 and args.seed == 0:
        args.output_path = Path.cwd().joinpath(args.output_path, "noisy_features",
                                               f"noise_{args.feature_noise}")
        if student:
            args.out_t_path = args.output_path
    if args.feature_aug_k > 0 and args.seed == 0:
        args.output_path = Path.cwd().joinpath(args.output_path, "aug_features",
                                               f"aug_hop_{args.feature_aug_k}")
        if student:
            args.student = f"GA{args.feature_aug_k}{args.student}"
        else:
            args.teacher = f"GA{args.feature_aug_k}{args.teacher}"

    model_name = args.student if student else args.teacher
    output_dir = _result_dir(args, args.output_path,
                             f"{args.teacher}_{args.student}" if student else args.teacher)
    args.output_dir = output_dir
    check_writable(output_dir, overwrite=False)
    if student:
        out_t_dir = _result_dir(args, args.out_t_path, args.teacher)
        check_readable(out_t_dir)
    logger = get_logger(output_dir.joinpath("log"), args.console_log, args.log_level)
    logger.info(f"output_dir: {output_dir}")
    if student:
        logger.info(f"out_t_dir: {out_t_dir}")

    g, labels, idx_train, idx_val, idx_test = load_data(
        args.dataset, args.data_path, split_idx=args.split_idx, seed=args.seed,
        labelrate_train=args.labelrate_train, labelrate_val=args.labelrate_val)
    logger.info(f"Total {g.number_of_nodes()} nodes.")
    logger.info(f"Total {g.number_of_edges()} edges.")
    feats = g.ndata["feat"]
    args.feat_dim = feats.shape[1]
    args.label_dim = labels.int().max().item() + 1
    if 0 < args.feature_noise <= 1:
        feats = (1 - args.feature_noise) * feats + args.feature_noise * torch.randn_like(feats)

    conf = {}
    if args.model_config_path is not None:
        conf = get_training_config(args.model_config_path, model_name, args.dataset)
    conf = dict(args.__dict__, **conf)  # YAML wins over the command line, as in the reference
    conf["device"] = device
    logger.info(f"conf: {conf}")

    model = Model(conf)
    optimizer = optim.Adam(model.parameters(), lr=conf["learning_rate"],
                           weight_decay=conf["weight_decay"])
    criterion = torch.nn.NLLLoss()
    evaluator = get_evaluator(conf["dataset"])
    if student:
        criterion_t = torch.nn.KLDivLoss(reduction="batchmean", log_target=True)
        out_t = _OUT_T_ON_DEVICE.get(str(out_t_dir))
        if out_t is None:
            out_t = load_out_t(out_t_dir)
        else:
            logger.info("teacher log-probabilities taken from device memory (same-process teacher run)")
        for name, idx in (("train", idx_train), ("val", idx_val), ("test", idx_test)):
            logger.debug(f"teacher score on {name} data: "
                         f"{evaluator(out_t[idx.to(out_t.device)], labels[idx].to(out_t.device))}")

    def propagate(x, graph):
        dev = device if device != "cpu" else x.device
        return feature_prop(x.to(dev), graph, args.feature_aug_k).to(x.device)

    loss_and_score = []
    if args.exp_setting == "tran":
        if args.feature_aug_k > 0:
            feats = propagate(feats, g)
        if student:
            indices = (idx_train, torch.cat([idx_train, idx_val, idx_test]), idx_val, idx_test)
            out, _, score_test = TE.distill_run_transductive(
                conf, model, feats, labels, out_t, indices, criterion, criterion_t, evaluator,
                optimizer, logger, loss_and_score)
        else:
            out, _, score_test = TE.run_transductive(
                conf, model, g, feats, labels, (idx_train, idx_val, idx_test), criterion, evaluator,
                optimizer, logger, loss_and_score)
        score_lst = [score_test]
    else:
        split = graph_split(idx_train, idx_val, idx_test, args.split_rate, args.seed)
        obs_idx_train, obs_idx_val, obs_idx_test, idx_obs, idx_test_ind = split
        if args.feature_aug_k > 0:  # the observed graph only propagates within itself
            obs_feats = propagate(feats[idx_obs], g.subgraph(idx_obs))
            feats = propagate(feats, g)
            feats[idx_obs] = obs_feats
        if student:
            indices = (obs_idx_train, torch.cat([obs_idx_train, obs_idx_val, obs_idx_test]),
                       obs_idx_val, obs_idx_test, idx_obs, idx_test_ind)
            out, _, s_tran, s_ind = TE.distill_run_inductive(
                conf, model, feats, labels, out_t, indices, criterion, criterion_t, evaluator,
                optimizer, logger, loss_and_score)
        else:
            out, _, s_tran, s_ind = TE.run_inductive(
                conf, model, g, feats, labels, split, criterion, evaluator, optimizer, logger,
                loss_and_score)
        score_lst = [s_tran, s_ind]

    logger.info(f"num_layers: {conf['num_layers']}. hidden_dim: {conf['hidden_dim']}. "
                f"dropout_ratio: {conf['dropout_ratio']}")
    logger.info(f"# params {sum(p.numel() for p in model.parameters())}")

    np.savez(output_dir.joinpath("out"), out.detach().cpu().numpy())  # out.npz, key arr_0
    if not student:
        _OUT_T_ON_DEVICE.clear()   # one teacher output at a time stays resident
        _OUT_T_ON_DEVICE[str(output_dir)] = out.detach()
    if args.save_results:
        np.savez(output_dir.joinpath("loss_and_score"), np.array(loss_and_score))
        torch.save(model.state_dict(), output_dir.joinpath("model.pth"))
    if args.exp_setting == "tran" and args.compute_min_cut:
        with open(output_dir.parent.joinpath("min_cut_loss"), "a+") as f:
            f.write(f"{compute_min_cut_loss(g, out) :.4f}\n")
    return score_lst


def distill_pipeline(teacher_argv, student_argv):
    """train_teacher.py then train_student.py in one process: the teacher's log-probabilities stay on
    the device between the two (out.npz is still written).  Returns (teacher line, student line)."""
    return main("teacher", teacher_argv), main("student", student_argv)


def main(role, argv=None):
    args = build_parser(role).parse_args(argv)
    if args.num_exp == 1:
        scores = run(args, role)
        score_str = "".join(f"{s : .4f}\t" for s in scores)
    elif args.num_exp > 1:
        runs = []
        for seed in range(args.num_exp):
            args.seed = seed
            runs.append(run(args, role))
        runs = np.array(runs)
        score_str = "".join([f"{s : .4f}\t" for s in runs.mean(axis=0)] +
                            [f"{s : .4f}\t" for s in runs.std(axis=0)])
    else:
        raise ValueError("--num_exp must be >= 1")
    with open(args.output_dir.parent.joinpath("exp_results"), "a+") as f:
        f.write(f"{score_str}\n")
    print(score_str)
    return score_str
