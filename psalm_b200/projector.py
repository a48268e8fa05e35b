"""`ResNetSwin` projector: res5 -> LLM image tokens (reference multimodal_projector/builder.py:326-375,
BasicBlock :85-111).  conv3x3 s2 (1024->2048)+BN+ReLU, conv3x3 (2048->2048) applied TWICE with the same
weights (builder.py:92-94), BN, 1x1 s2 shortcut + BN, add, ReLU, flatten, fc.  Eval-mode BatchNorm is
folded into the adjacent convolution at weight-preparation time; convolutions are library (cuDNN)
calls on channels-last tensors."""
import torch
import torch.nn.functional as F


def _fold(sd, bn, eps=1e-5):
    scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + eps)
    shift = sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale
    return scale, shift


class ResNetSwin:
    def __init__(self, sd, prefix="model.mm_projector.", dtype=torch.bfloat16, device="cuda"):
        self.dtype, self.device = dtype, device
        b = prefix + "layer1.0."
        cl = lambda t: t.to(device=device, dtype=dtype).contiguous(memory_format=torch.channels_last)  # noqa: E731
        cv = lambda t: t.to(device=device, dtype=dtype).contiguous()  # noqa: E731
        s1, t1 = _fold(sd, b + "bn1")
        s2, t2 = _fold(sd, b + "bn2")
        sd_, td = _fold(sd, b + "downsample.1")
        self.w1, self.b1 = cl(sd[b + "conv1.weight"].float() * s1.view(-1, 1, 1, 1)), cv(t1)
        self.w2 = cl(sd[b + "conv2.weight"])
        self.w2f, self.b2 = cl(sd[b + "conv2.weight"].float() * s2.view(-1, 1, 1, 1)), cv(t2)
        self.wd, self.bd = cl(sd[b + "downsample.0.weight"].float() * sd_.view(-1, 1, 1, 1)), cv(td)
        self.fcw, self.fcb = cv(sd[prefix + "fc.weight"]), cv(sd[prefix + "fc.bias"])

    def __call__(self, res5):
        return self.forward(res5)

    def forward(self, res5):
        """res5 NCHW (any strides) [B,1024,h,w] -> [B, ceil(h/2)*ceil(w/2), hidden]."""
        x = res5.to(dtype=self.dtype).contiguous(memory_format=torch.channels_last)
        out = F.relu(F.conv2d(x, self.w1, self.b1, stride=2, padding=1))
        out = F.conv2d(out, self.w2, None, padding=1)
        out = F.conv2d(out, self.w2f, self.b2, padding=1)
        out = F.relu(out + F.conv2d(x, self.wd, self.bd, stride=2))
        B, C, h, w = out.shape
        out = out.permute(0, 2, 3, 1).reshape(B, h * w, C)
        return F.linear(out, self.fcw, self.fcb)
