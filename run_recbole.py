"""Same command line as the reference's run_recbole.py (:14-27):
    python run_recbole.py --model=ACSASRec --dataset=ml-100k --config_files="config/config_t.yaml config/x.yaml" [--key=value ...]
"""
import argparse
import sys

import ac_tsr_b200 as A

if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--model', '-m', type=str, default='ACSASRec', help='name of models')
    parser.add_argument('--dataset', '-d', type=str, default='ml-100k', help='name of datasets')
    parser.add_argument('--config_files', type=str, default=None, help='config files')
    args, rest = parser.parse_known_args()
    config_file_list = args.config_files.strip().split(' ') if args.config_files else None
    A.quick_start.run_recbole(model=args.model, dataset=args.dataset, config_file_list=config_file_list, cmd_args=rest)
