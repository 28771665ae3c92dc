"""RGB feature pyramid and small conv wrappers (mirror of src/modules/basicConv.py:
createCNNs :6-20, Conv1d :63-85).  Module indices inside the Sequential -- hence the
state_dict keys `RGB_net1.{0,1,4,5,...}` -- follow the reference: block i owns slots
4i (conv), 4i+1 (BatchNorm2d), 4i+2 (LeakyReLU 0.1), 4i+3 (MaxPool 3x3)."""
import torch.nn as nn


def createCNNs(in_channel, channels, strides):
    layers = nn.Sequential()
    last = in_channel
    for i, (out_channel, stride) in enumerate(zip(channels, strides)):
        layers.add_module(str(4 * i), nn.Conv2d(last, out_channel, kernel_size=3, stride=1, padding=1, bias=True))
        layers.add_module(str(4 * i + 1), nn.BatchNorm2d(out_channel))
        layers.add_module(str(4 * i + 2), nn.LeakyReLU(negative_slope=0.1))
        layers.add_module(str(4 * i + 3), nn.MaxPool2d(3, stride=stride, padding=1))
        last = out_channel
    return layers


class Conv1d(nn.Module):
    """(B,N,C) -> (B,N,C'): Conv1d(k=1) [+ BatchNorm1d] [+ (Leaky)ReLU] under `composed_module`."""

    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=0, use_activation=True,
                 use_leaky=True, bn=False):
        super().__init__()
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, kernel_size
        if use_activation:
            act = nn.LeakyReLU(0.1, inplace=True) if use_leaky else nn.ReLU(inplace=True)
        else:
            act = nn.Identity()
        self.composed_module = nn.Sequential(
            nn.Conv1d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=padding),
            nn.BatchNorm1d(out_channels) if bn else nn.Identity(),
            act)

    def forward(self, x):
        return self.composed_module(x.permute(0, 2, 1)).permute(0, 2, 1)
