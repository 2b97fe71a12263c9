"""ORACLE (image path) — test infrastructure only.

Plain-torch CPU restatement of the Omniglot model on the aggressive inner loop:
ResNetEncoderV2 (modules/encoders/enc_resnet_v2.py:93-126) and PixelCNNDecoderV2.reconstruct_error
(modules/decoders/dec_pixelcnn_v2.py:123-195), written with explicit functional ops on a flat parameter dict
whose keys are the reference `state_dict` keys.  Parity status: PINNED by
`oracle/validate_image_against_reference.py` (imports the unmodified reference here; fixtures in tests/golden/).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
KS_MAIN = [7, 7, 7, 7, 7, 5, 5, 5, 5, 3, 3, 3, 3]          # dec_pixelcnn_v2.py:140 ('large')
KS_DIRECT = KS_MAIN[1:-1]                                    # dec_pixelcnn_v2.py:102-104


def bn_train(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d in train(): batch statistics, biased variance."""
    mean = x.mean(dim=(0, 2, 3), keepdim=True)
    var = x.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def bn(p: Dict[str, Tensor], pre: str, x: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d named `pre` (keys pre + weight / bias / running_mean / running_var).  train(): batch statistics
    (bn_train); eval(): the running statistics — selected by p["__eval__"] (absent = train, the aggressive-loop case)."""
    if not p.get("__eval__", False):
        return bn_train(x, p[pre + "weight"], p[pre + "bias"], eps)
    m, v = p[pre + "running_mean"].view(1, -1, 1, 1), p[pre + "running_var"].view(1, -1, 1, 1)
    return (x - m) / torch.sqrt(v + eps) * p[pre + "weight"].view(1, -1, 1, 1) + p[pre + "bias"].view(1, -1, 1, 1)


def masked_weight(weight: Tensor, mask: Tensor) -> Tensor:
    """MaskedConv2d.forward multiplies `weight.data` by the mask IN PLACE (dec_pixelcnn_v2.py:29) and then convolves with
    the parameter itself: the value is w*mask but autograd sees a plain convolution, so the gradient of EVERY tap (masked
    ones included) is non-zero.  Same value and same gradient here, without mutating the input."""
    return weight + (weight * mask - weight).detach()


def conv_mask(weight: Tensor, mask_type: str, masked_channels: int) -> Tensor:
    """MaskedConv2d mask (dec_pixelcnn_v2.py:16-20)."""
    m = torch.ones_like(weight)
    kH, kW = weight.shape[2], weight.shape[3]
    m[:, :masked_channels, kH // 2, kW // 2 + (1 if mask_type == "B" else 0):] = 0
    m[:, :masked_channels, kH // 2 + 1:] = 0
    return m


# ------------------------------------------------------------------------------------------------------
# parameter spec (reference state_dict keys and shapes) + deterministic test parameters
# ------------------------------------------------------------------------------------------------------
def _bn(spec, pre, c):
    spec[pre + "weight"] = (c,)
    spec[pre + "bias"] = (c,)
    spec[pre + "running_mean"] = (c,)
    spec[pre + "running_var"] = (c,)
    spec[pre + "num_batches_tracked"] = ()


def image_param_spec(nz: int, fm: int = 4) -> Dict[str, tuple]:
    """Every state_dict entry of VAE(ResNetEncoderV2, PixelCNNDecoderV2) in reference order
    (enc_resnet_v2.py:27-45,93-107; dec_pixelcnn_v2.py:12-20,32-48,65-75,88-106,123-152)."""
    spec: Dict[str, tuple] = {}
    cin = 1
    for i in range(3):
        pre = "encoder.main.0.main.%d." % i
        spec[pre + "conv1.weight"] = (64, cin, 3, 3)
        _bn(spec, pre + "bn1.", 64)
        spec[pre + "conv2.weight"] = (64, 64, 3, 3)
        _bn(spec, pre + "bn2.", 64)
        spec[pre + "downsample.0.weight"] = (64, cin, 1, 1)
        _bn(spec, pre + "downsample.1.", 64)
        cin = 64
    spec["encoder.main.1.weight"] = (512, 64, 4, 4)
    _bn(spec, "encoder.main.2.", 512)
    spec["encoder.linear.weight"] = (2 * nz, 512)
    spec["encoder.linear.bias"] = (2 * nz,)
    spec["decoder.z_transform.0.weight"] = (fm * 784, nz)
    spec["decoder.z_transform.0.bias"] = (fm * 784,)

    def block(pre, k):
        spec[pre + "main.0.weight"] = (32, 64, 1, 1)
        _bn(spec, pre + "main.1.", 32)
        spec[pre + "main.3.weight"] = (32, 32, k, k)
        spec[pre + "main.3.mask"] = (32, 32, k, k)
        _bn(spec, pre + "main.4.", 32)
        spec[pre + "main.6.weight"] = (64, 32, 1, 1)
        _bn(spec, pre + "main.7.", 64)
    pre = "decoder.main.0."
    spec[pre + "main.0.main.0.weight"] = (64, 1 + fm, 7, 7)
    spec[pre + "main.0.main.0.mask"] = (64, 1 + fm, 7, 7)
    _bn(spec, pre + "main.0.main.1.", 64)
    for i in range(1, len(KS_MAIN)):
        block(pre + "main.%d." % i, KS_MAIN[i])
    for i, k in enumerate(KS_DIRECT):
        block(pre + "direct_connects.%d." % i, k)
    spec["decoder.main.1.weight"] = (64, 64, 1, 1)
    _bn(spec, "decoder.main.2.", 64)
    spec["decoder.main.4.weight"] = (1, 64, 1, 1)
    return spec


def init_image_params(nz: int, seed: int = 0, fm: int = 4) -> Dict[str, Tensor]:
    """Deterministic, 'trained-like' test parameters (conv N(0, sqrt(2/(k k out))) as the reference initialisers,
    but BatchNorm gamma/beta and biases randomised so that they are exercised).  Fixture helper."""
    g = torch.Generator().manual_seed(seed)
    p: Dict[str, Tensor] = {}
    for k, shp in image_param_spec(nz, fm).items():
        if k.endswith("num_batches_tracked"):
            p[k] = torch.zeros((), dtype=torch.int64)
        elif k.endswith("running_mean"):
            p[k] = torch.zeros(shp)
        elif k.endswith("running_var"):
            p[k] = torch.ones(shp)
        elif k.endswith(".mask"):
            continue
        elif len(shp) == 4:
            std = (2.0 / (shp[2] * shp[3] * shp[0])) ** 0.5
            p[k] = torch.randn(shp, generator=g) * std
        elif len(shp) == 2:
            a = (6.0 / (shp[0] + shp[1])) ** 0.5
            p[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
        elif k.endswith("bias") and ("linear" in k or "z_transform" in k):
            p[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
        elif k.endswith("weight"):      # BatchNorm gamma
            p[k] = 0.5 + torch.rand(shp, generator=g)
        else:                           # BatchNorm beta
            p[k] = (torch.rand(shp, generator=g) * 2 - 1) * 0.2
    for k, shp in image_param_spec(nz, fm).items():
        if k.endswith(".mask"):
            w = p[k.replace(".mask", ".weight")]
            p[k] = conv_mask(w, "A" if k.endswith("main.0.main.0.mask") else "B", 1 if k.endswith("main.0.main.0.mask") else shp[1])
    return {k: p[k] for k in image_param_spec(nz, fm)}


def make_image_batch(B: int, seed: int = 1234) -> Tensor:
    """Dynamically binarised synthetic images (image.py:287)."""
    g = torch.Generator().manual_seed(seed)
    return torch.bernoulli(torch.rand(B, 1, 28, 28, generator=g), generator=g)


# ------------------------------------------------------------------------------------------------------
# encoder — enc_resnet_v2.py
# ------------------------------------------------------------------------------------------------------
def resnet_block(p: Dict[str, Tensor], pre: str, x: Tensor, stride: int) -> Tensor:
    """ResNetBlock.forward (enc_resnet_v2.py:55-71); downsample branch always present here (stride 2)."""
    res = bn(p, pre + "downsample.1.", F.conv2d(x, p[pre + "downsample.0.weight"], stride=stride))
    out = F.elu(bn(p, pre + "bn1.", F.conv2d(x, p[pre + "conv1.weight"], stride=stride, padding=1)))
    out = bn(p, pre + "bn2.", F.conv2d(out, p[pre + "conv2.weight"], padding=1))
    return F.elu(out + res)


def encoder_forward(p: Dict[str, Tensor], x: Tensor) -> Tuple[Tensor, Tensor]:
    """ResNetEncoderV2.forward (enc_resnet_v2.py:120-126) -> (mu, logvar)."""
    h = x
    for i in range(3):
        h = resnet_block(p, "encoder.main.0.main.%d." % i, h, 2)
    h = F.elu(bn(p, "encoder.main.2.", F.conv2d(h, p["encoder.main.1.weight"])))
    out = h.view(h.shape[0], -1) @ p["encoder.linear.weight"].t() + p["encoder.linear.bias"]
    nz = out.shape[1] // 2
    return out[:, :nz], out[:, nz:]


# ------------------------------------------------------------------------------------------------------
# decoder — dec_pixelcnn_v2.py
# ------------------------------------------------------------------------------------------------------
def pixelcnn_block(p: Dict[str, Tensor], pre: str, x: Tensor, k: int) -> Tensor:
    """PixelCNNBlock.forward (dec_pixelcnn_v2.py:32-62)."""
    c = x.shape[1] // 2
    h = F.elu(bn(p, pre + "main.1.", F.conv2d(x, p[pre + "main.0.weight"])))
    w = masked_weight(p[pre + "main.3.weight"], conv_mask(p[pre + "main.3.weight"], "B", c))
    h = F.elu(bn(p, pre + "main.4.", F.conv2d(h, w, padding=k // 2)))
    h = bn(p, pre + "main.7.", F.conv2d(h, p[pre + "main.6.weight"]))
    return F.elu(h + x)


def pixelcnn_forward(p: Dict[str, Tensor], inp: Tensor) -> Tensor:
    """PixelCNN.forward wiring (dec_pixelcnn_v2.py:108-121) + head (145-152) -> probabilities [N,1,28,28]."""
    pre = "decoder.main.0."
    directs: List[Tensor] = []
    h = inp
    for i, k in enumerate(KS_MAIN):
        if i > 2:
            d_in = directs.pop(0)
            h = h + pixelcnn_block(p, pre + "direct_connects.%d." % (i - 3), d_in, KS_DIRECT[i - 3])
        if i == 0:                                                            # MaskABlock (65-85)
            w = masked_weight(p[pre + "main.0.main.0.weight"], conv_mask(p[pre + "main.0.main.0.weight"], "A", 1))
            h = F.elu(bn(p, pre + "main.0.main.1.", F.conv2d(h, w, padding=k // 2)))
        else:
            h = pixelcnn_block(p, pre + "main.%d." % i, h, k)
        directs.append(h)
    h = h + pixelcnn_block(p, pre + "direct_connects.%d." % (len(KS_DIRECT) - 1), directs.pop(0), KS_DIRECT[-1])
    h = F.elu(bn(p, "decoder.main.2.", F.conv2d(h, p["decoder.main.1.weight"])))
    return torch.sigmoid(F.conv2d(h, p["decoder.main.4.weight"]))


def decoder_reconstruct_error(p: Dict[str, Tensor], x: Tensor, z: Tensor, fm: int = 4) -> Tensor:
    """PixelCNNDecoderV2.reconstruct_error (dec_pixelcnn_v2.py:172-195): x [B,1,28,28], z [B,ns,nz] -> [B,ns]."""
    B, ns, _ = z.shape
    zt = (z @ p["decoder.z_transform.0.weight"].t() + p["decoder.z_transform.0.bias"]).view(B, ns, fm, 28, 28)
    img = torch.cat([x.unsqueeze(1).expand(B, ns, *x.shape[1:]), zt], dim=2).reshape(B * ns, 1 + fm, 28, 28)
    rec = pixelcnn_forward(p, img).view(B, ns, -1)
    xf = x.view(B, -1).unsqueeze(1)
    bce = (rec + 1e-12).log() * xf + (1.0 - rec + 1e-12).log() * (1.0 - xf)
    return -bce.sum(dim=2)


def vae_loss(p: Dict[str, Tensor], x: Tensor, kl_weight: float, eps: Tensor, fm: int = 4, training: bool = True):
    """VAE.loss (modules/vae.py:79-98) for the image model; eps [B,ns,nz].  training=False: eval() — every BatchNorm
    normalises with its running statistics (image.py test() / calc_mi / sampling)."""
    if not training:
        p = dict(p)
        p["__eval__"] = True
    mu, logvar = encoder_forward(p, x)
    z = mu.unsqueeze(1) + eps * (0.5 * logvar).exp().unsqueeze(1)
    KL = 0.5 * (mu.pow(2) + logvar.exp() - logvar - 1).sum(dim=1)
    rec = decoder_reconstruct_error(p, x, z, fm).mean(dim=1)
    return rec + kl_weight * KL, rec, KL
