"""
Command-line + HOCON config ingestion with the reference's flag set (src/util/args.py:9-112)
so that its eval / train scripts can call `util.args.parse_args(...)` unchanged.  Uses the
real pyhocon when installed, otherwise the in-repo subset reader (util/hocon.py).
"""
import argparse
import os

try:  # pragma: no cover - pyhocon is not in this image
    from pyhocon import ConfigFactory
except ImportError:
    from .hocon import ConfigFactory

PROJECT_ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))


def build_parser(default_expname, default_num_epochs, default_lr, default_gamma, default_ray_batch_size):
    p = argparse.ArgumentParser()
    p.add_argument("--conf", "-c", type=str, default=None)
    p.add_argument("--resume", "-r", action="store_true", help="continue training")
    p.add_argument("--gpu_id", type=str, default="0", help="GPU(s) to use, space delimited")
    p.add_argument("--name", "-n", type=str, default=default_expname, help="experiment name")
    p.add_argument("--dataset_format", "-F", type=str, default=None,
                   help="Dataset format, multi_obj | dvr | dvr_gen | dvr_dtu | srn")
    p.add_argument("--exp_group_name", "-G", type=str, default=None, help="experiment group")
    p.add_argument("--logs_path", type=str, default="logs", help="logs output directory")
    p.add_argument("--checkpoints_path", type=str, default="checkpoints", help="checkpoints directory")
    p.add_argument("--visual_path", type=str, default="visuals", help="visualization directory")
    p.add_argument("--epochs", type=int, default=default_num_epochs, help="number of epochs to train for")
    p.add_argument("--lr", type=float, default=default_lr, help="learning rate")
    p.add_argument("--gamma", type=float, default=default_gamma, help="learning rate decay factor")
    p.add_argument("--datadir", "-D", type=str, default=None, help="Dataset directory")
    p.add_argument("--ray_batch_size", "-R", type=int, default=default_ray_batch_size, help="Ray batch size")
    return p


def parse_args(callback=None, training=False, default_conf="conf/default_mv.conf", default_expname="example",
               default_data_format="dvr", default_num_epochs=10000000, default_lr=1e-4, default_gamma=1.00,
               default_datadir="data", default_ray_batch_size=50000, argv=None):
    parser = build_parser(default_expname, default_num_epochs, default_lr, default_gamma, default_ray_batch_size)
    if callback is not None:
        parser = callback(parser)
    args = parser.parse_args(argv)

    if args.exp_group_name is not None:
        for attr in ("logs_path", "checkpoints_path", "visual_path"):
            setattr(args, attr, os.path.join(getattr(args, attr), args.exp_group_name))
    os.makedirs(os.path.join(args.checkpoints_path, args.name), exist_ok=True)
    os.makedirs(os.path.join(args.visual_path, args.name), exist_ok=True)

    expconf_path = os.path.join(PROJECT_ROOT, "expconf.conf")
    expconf = ConfigFactory.parse_file(expconf_path)
    if args.conf is None:
        args.conf = expconf.get_string("config." + args.name, default_conf)
    if args.datadir is None:
        args.datadir = expconf.get_string("datadir." + args.name, default_datadir)
    conf_path = args.conf if os.path.exists(args.conf) else os.path.join(PROJECT_ROOT, args.conf)
    conf = ConfigFactory.parse_file(conf_path)
    if args.dataset_format is None:
        args.dataset_format = conf.get_string("data.format", default_data_format)
    args.gpu_id = list(map(int, args.gpu_id.split()))

    print("EXPERIMENT NAME:", args.name)
    if training:
        print("CONTINUE?", "yes" if args.resume else "no")
    print("* Config file:", args.conf)
    print("* Dataset format:", args.dataset_format)
    print("* Dataset location:", args.datadir)
    return args, conf
