"""Named teacher models for ``python -m paddle_edl.distill.teacher_server --model NAME``.
``build(name) -> (module, feed_names, fetch_names, feed_shapes)``; the fetch name of the image
teachers is ``score`` (softmax probabilities), matching the reference's serving model
(example/distill/resnet/train_with_fleet.py:446)."""
import torch
import torch.nn as nn


class _Probs(nn.Module):
    def __init__(self, net, channels_last=True):
        super().__init__()
        self.net, self.channels_last = net, channels_last

    @torch.no_grad()
    def forward(self, x):
        if x.dim() == 4 and self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        return torch.softmax(self.net(x).float(), -1)


def build(name):
    name = name.lower()
    if name in ("resnext101_32x16d", "resnext101_32x16d_wsl"):
        from .resnext import ResNeXt101_32x16d
        return _Probs(ResNeXt101_32x16d()), ["image"], ["score"], {"image": [3, 224, 224]}
    if name == "resnext50_32x4d":
        from .resnext import ResNeXt50_32x4d
        return _Probs(ResNeXt50_32x4d()), ["image"], ["score"], {"image": [3, 224, 224]}
    if name == "resnext_tiny":
        from .resnext import ResNeXt_tiny
        return _Probs(ResNeXt_tiny()), ["image"], ["score"], {"image": [3, 64, 64]}
    if name == "mnist_cnn":
        from .small import MnistTeacher
        return _Probs(MnistTeacher(), channels_last=False), ["img"], ["fc_0.tmp_2"], {"img": [1, 28, 28]}
    raise ValueError("unknown teacher model %r" % name)
