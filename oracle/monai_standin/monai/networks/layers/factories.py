"""monai.networks.layers.factories.Conv restated: Conv["conv", 2] -> nn.Conv2d, Conv["convtrans", 2] -> nn.ConvTranspose2d."""
import torch.nn as nn


class _ConvFactory:
    CONV = "conv"
    CONVTRANS = "convtrans"
    _table = {
        ("conv", 1): nn.Conv1d, ("conv", 2): nn.Conv2d, ("conv", 3): nn.Conv3d,
        ("convtrans", 1): nn.ConvTranspose1d, ("convtrans", 2): nn.ConvTranspose2d, ("convtrans", 3): nn.ConvTranspose3d,
    }

    def __getitem__(self, key):
        name, dim = key
        return self._table[(str(name).lower(), int(dim))]


Conv = _ConvFactory()
