"""`run_recbole(model='ACSASRec', dataset=..., config_file_list=..., config_dict=...)` -- the reference's entry point
(recbole/quick_start/quick_start.py:21-77) over this package's Config, SequentialDataset, loaders, ACSASRec and ACSASRecTrainer.
Same control flow, same return dictionary; the model and trainer are resolved by name exactly like utils.get_model /
utils.get_trainer would (only AC-SASRec lives here)."""
import logging
import random

import numpy as np
import torch

from .acsasrec import ACSASRec
from .acbert4rec import AcBERT4Rec
from .acssept import ACSSEPT
from .actisasrec import ACTiSASRec
from .compat import Config
from .dataset import create_dataset, data_preparation
from .trainer import ACSASRecTrainer, AcBERT4RecTrainer, ACSSEPTTrainer, ACTiSASRecTrainer

_MODELS = {'ACSASRec': (ACSASRec, ACSASRecTrainer), 'AcBERT4Rec': (AcBERT4Rec, AcBERT4RecTrainer),
           'ACSSEPT': (ACSSEPT, ACSSEPTTrainer), 'ACTiSASRec': (ACTiSASRec, ACTiSASRecTrainer)}


def init_seed(seed, reproducibility):
    """utils/utils.py:129-147"""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    torch.backends.cudnn.benchmark = not reproducibility
    torch.backends.cudnn.deterministic = bool(reproducibility)


def get_model(name):
    if name not in _MODELS:
        raise ValueError('`model_name` [%s] is not the name of an existing model.' % name)         # utils.py:72-73
    return _MODELS[name][0]


def get_trainer(model_type, name):
    return _MODELS[name][1]


def run_recbole(model=None, dataset=None, config_file_list=None, config_dict=None, saved=True, cmd_args=None):
    config = Config(model=model, dataset=dataset, config_file_list=config_file_list, config_dict=config_dict, cmd_args=cmd_args)
    init_seed(config['seed'], config['reproducibility'])
    logging.basicConfig(level=getattr(logging, str(config['state'] or 'INFO').upper(), logging.INFO))
    logger = logging.getLogger()
    logger.info(config)
    ds = create_dataset(config)
    logger.info(ds)
    train_data, valid_data, test_data = data_preparation(config, ds)
    init_seed(config['seed'], config['reproducibility'])
    net = get_model(config['model'])(config, train_data.dataset).to(config['device'])
    logger.info(net)
    trainer = get_trainer(config['MODEL_TYPE'], config['model'])(config, net)
    best_valid_score, best_valid_result = trainer.fit(train_data, valid_data, saved=saved, show_progress=config['show_progress'])
    test_result = trainer.evaluate(test_data, load_best_model=saved, show_progress=config['show_progress'])
    logger.info('best valid : %s' % best_valid_result)
    logger.info('test result: %s' % test_result)
    return {'best_valid_score': best_valid_score, 'valid_score_bigger': config['valid_metric_bigger'],
            'best_valid_result': best_valid_result, 'test_result': test_result}


def objective_function(config_dict=None, config_file_list=None, saved=True):
    """quick_start.py:80-107: the hyper-tuning objective (same training run, quiet logging)."""
    config = Config(config_dict=config_dict, config_file_list=config_file_list)
    init_seed(config['seed'], config['reproducibility'])
    logging.basicConfig(level=logging.ERROR)
    ds = create_dataset(config)
    train_data, valid_data, test_data = data_preparation(config, ds)
    init_seed(config['seed'], config['reproducibility'])
    net = get_model(config['model'])(config, train_data.dataset).to(config['device'])
    trainer = get_trainer(config['MODEL_TYPE'], config['model'])(config, net)
    best_valid_score, best_valid_result = trainer.fit(train_data, valid_data, verbose=False, saved=saved)
    test_result = trainer.evaluate(test_data, load_best_model=saved)
    return {'best_valid_score': best_valid_score, 'valid_score_bigger': config['valid_metric_bigger'],
            'best_valid_result': best_valid_result, 'test_result': test_result}


def load_data_and_model(model_file):
    """quick_start.py:110-146: rebuild config, dataset, loaders and the trained model from a checkpoint written by the trainer
    (the checkpoint pickles the Config, hence weights_only=False -- SURVEY §8c)."""
    checkpoint = torch.load(model_file, map_location='cpu', weights_only=False)
    config = checkpoint['config']
    init_seed(config['seed'], config['reproducibility'])
    ds = create_dataset(config)
    train_data, valid_data, test_data = data_preparation(config, ds)
    init_seed(config['seed'], config['reproducibility'])
    net = get_model(config['model'])(config, train_data.dataset).to(config['device'])
    net.load_state_dict(checkpoint['state_dict'])
    net.load_other_parameter(checkpoint.get('other_parameter'))
    return config, net, ds, train_data, valid_data, test_data
