"""Drop-in for the reference's ``codes/network`` package (network/__init__.py:7-24): put this package's
parent directory on ``sys.path`` in place of the reference's and ``from network import build_model,
build_loss`` resolves to the B200-native implementation."""
from .loss import losswrapper
from .model_nefnet import Model_nefnet
from .model_nefnet2 import Model_nefnet2  # constructed directly, like the reference's class (build_model never returns it)


def build_model(cfg):
    """network/__init__.py:7-12"""
    model_name = cfg.MODEL.model
    if model_name == 'model_nefnet':
        return Model_nefnet(theta_encoder_len=cfg.MODEL.theta_L, lead_num=cfg.DATA.lead_num)
    raise ValueError('build model: model name error')


def build_loss(cfg):
    """network/__init__.py:15-24.  Only the Standin-Learning wrapper ('v1') is on the hot path; the
    stock 'ce' / 'mse' criteria are returned as the PyTorch modules the reference returns."""
    loss_name = cfg.MODEL.loss
    if loss_name == 'v1':
        return losswrapper
    if loss_name == 'ce':
        from torch.nn import CrossEntropyLoss
        return CrossEntropyLoss()
    if loss_name == 'mse':
        from torch.nn import MSELoss
        return MSELoss()
    raise ValueError('build loss: loss name error')
