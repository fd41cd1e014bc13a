"""Drop-in for reference ``src/models/cnn/base.py`` (`CNN`, `ContextGating`): the PMAM / DASM convolutional branch.

Module / parameter names match the reference (``cnn.conv{i}``, ``cnn.batchnorm{i}``, ``cnn.cg{i}.linear``, ...), so checkpoints
load unchanged.  The arithmetic runs channels-last through libt4s: 3x3 convolutions and the gating Linear as tcgen05 GEMMs on
[pixels, channels], BatchNorm / gate / dropout / pooling as fused row kernels (`functional.conv3x3`, `batch_norm`,
`context_gate`, `avg_pool`).  Only the configuration the shipped PMAM / DASM YAMLs select is provided: kernel 3, stride 1,
padding 1, BatchNorm, ContextGating activation.
"""
import torch.nn as nn

from ... import functional as F


class ContextGating(nn.Module):
    def __init__(self, input_num):
        super().__init__()
        self.sigmoid = nn.Sigmoid()
        self.linear = nn.Linear(input_num, input_num)

    def forward_cl(self, x, dropout_p=0.0):
        """x channels-last [..., C] -> x * sigmoid(linear(x)), optional dropout fused into the product."""
        return F.context_gate(x, F.linear(x, self.linear.weight, self.linear.bias), dropout_p)

    def forward(self, x):
        """Reference layout [B, C, H, W]."""
        return self.forward_cl(F.to_act(x.permute(0, 2, 3, 1).contiguous())).permute(0, 3, 1, 2)


class CNN(nn.Module):
    def __init__(self, n_in_channel, activation="Relu", conv_dropout=0, kernel_size=[3, 3, 3], padding=[1, 1, 1], stride=[1, 1, 1],
                 nb_filters=[64, 64, 64], pooling=[(1, 4), (1, 4), (1, 4)], normalization="batch"):
        super().__init__()
        assert len(kernel_size) == len(padding) == len(stride) == len(nb_filters) == len(pooling)
        if activation.lower() != "cg" or normalization != "batch":
            raise NotImplementedError("CNN: only activation='cg' with normalization='batch' (the PMAM / DASM configs) is on the B200 path")
        if any(k != 3 for k in kernel_size) or any(p != 1 for p in padding) or any(s != 1 for s in stride):
            raise NotImplementedError("CNN: 3x3 / stride 1 / padding 1 convolutions only")
        self.nb_filters = nb_filters
        self.pooling_cfg = [tuple(p) for p in pooling]
        self.conv_dropout = conv_dropout
        cnn = nn.Sequential()
        for i in range(len(nb_filters)):
            n_in = n_in_channel if i == 0 else nb_filters[i - 1]
            cnn.add_module("conv{0}".format(i), nn.Conv2d(n_in, nb_filters[i], 3, 1, 1))
            cnn.add_module("batchnorm{0}".format(i), nn.BatchNorm2d(nb_filters[i], eps=0.001, momentum=0.99))
            cnn.add_module("cg{0}".format(i), ContextGating(nb_filters[i]))
            if conv_dropout is not None:
                cnn.add_module("dropout{0}".format(i), nn.Dropout(conv_dropout))
            cnn.add_module("pooling{0}".format(i), nn.AvgPool2d(self.pooling_cfg[i]))
        self.cnn = cnn

    def forward_cl(self, x, mel_layout=False):
        """Channels-last forward.  x [B, T, F, C_in], or with `mel_layout` the log-mel image [B, F, T] itself (read in place as
        [B, T, F, 1]).  Returns [B, T', F', C_out]."""
        for i in range(len(self.nb_filters)):
            conv, bn, cg = getattr(self.cnn, f"conv{i}"), getattr(self.cnn, f"batchnorm{i}"), getattr(self.cnn, f"cg{i}")
            x = F.conv3x3(x, conv.weight, conv.bias, mel_layout=mel_layout and i == 0)
            training = self.training and bn.training
            x = F.batch_norm(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps, bn.momentum, training)
            if training and bn.num_batches_tracked is not None:
                bn.num_batches_tracked.add_(1)
            p = float(self.conv_dropout) if (self.training and self.conv_dropout) else 0.0
            x = cg.forward_cl(x, dropout_p=p)
            x = F.avg_pool(x, *self.pooling_cfg[i])
        return x

    def forward(self, x):
        """Reference contract: x [B, C_in, T, F] -> [B, C_out, T', F']."""
        return self.forward_cl(F.to_act(x.permute(0, 2, 3, 1).contiguous())).permute(0, 3, 1, 2)
