"""CLI with the reference's flags (lstm/test_net.py:19-38): ``python -m lstm_ctc_ocr_b200.lstm.test_net --network=LSTM_test
--cfg=./lstm/lstm.yml --testDir ./data/val`` (what test.sh runs)."""
import argparse
import pprint
import sys


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Test a lstm network")
    p.add_argument("--gpu", dest="gpu_id", default=0, type=int)
    p.add_argument("--cfg", dest="cfg_file", default=None, type=str)
    p.add_argument("--network", dest="network_name", default="LSTM_test", type=str)
    p.add_argument("--testDir", dest="test_dir", default="./data/val", type=str)
    p.add_argument("--weights", dest="pretrained_model", default=None, type=str)
    p.add_argument("--set", dest="set_cfgs", default=None, nargs=argparse.REMAINDER)
    p.add_argument("--restore", dest="restore", default=1, type=int)
    if argv is None and len(sys.argv) == 1:
        p.print_help()
    return p.parse_args(argv)


def main(argv=None):
    from ..lib.lstm.config import AttrDict, cfg, cfg_from_file, cfg_from_list, get_log_dir, get_output_dir
    from ..lib.lstm.test import test_net
    from ..lib.networks.factory import get_network
    args = parse_args(argv)
    if args.cfg_file is not None:
        cfg_from_file(args.cfg_file)
    if args.set_cfgs is not None:
        cfg_from_list(args.set_cfgs)
    pprint.pprint(cfg)
    imgdb = AttrDict(name="lstm_test")
    output_dir = get_output_dir(imgdb, None)
    log_dir = get_log_dir(imgdb)
    network = get_network(args.network_name)
    test_net(network, imgdb, args.test_dir, output_dir, log_dir, pretrained_model=args.pretrained_model, restore=bool(args.restore))


if __name__ == "__main__":
    main()
