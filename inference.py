"""Drop-in for the reference's `inference.py`: design / optimize / trajectory sampling over a list of
preprocessed `.npz` complexes (same command line; see abx_b200/cli.py for the few additions).

    python inference.py --model abx_diffab.ckpt --model_features abx_b200/config/config_data_feature.json \\
        --model_config abx_b200/config/config_model.json --name_idx test_data/diffab_test.idx \\
        --data_dir <npz dir> --output_dir out --mode design --num_samples 16
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from abx_b200 import cli  # noqa: E402


def load_batches(args):
    from abx_b200.data import dataset
    with open(args.name_idx) as f:
        names = [x.strip() for x in f if x.strip()]
    return dataset.load(args.data_dir, names, feats=None, batch_size=args.batch_size)


if __name__ == '__main__':
    cli.main(cli.build_parser(single_pdb=False).parse_args(), load_batches)
