"""CPU restatement of the ERFNet encoder/decoder forward (TEST INFRASTRUCTURE --
see oracle/__init__.py).  Plain torch functional ops on a parameter dict keyed
like the reference ``state_dict`` (minus the ``net.`` prefix); backward is
torch autograd.  dtype follows the parameters (float32 = reference arithmetic,
float64 = arbiter).

Restates Backprojection_Loss/Networks/ERFNet.py:
  DownsamplerBlock :11-22, non_bottleneck_1d :25-60, Encoder :63-95,
  UpsamplerBlock :98-107, Decoder :109-161, Net :164-176.
BatchNorm: eps 1e-3 (:17,33,39,102), training-mode batch statistics unless
``training=False`` (then ``running_mean/var`` are read from the dict).
Dropout2d (:41,57-58): ``drop_masks`` maps block prefix -> [B,C] keep-mask already
scaled by 1/(1-p); absent = no dropout (p=0 / eval).
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-3

ENC_NB = [("encoder.layers.%d" % i, 1) for i in range(1, 6)]
ENC_NB += [("encoder.layers.%d" % (7 + i), d) for i, d in enumerate([2, 4, 8, 16, 2, 4, 8, 16])]


def _bn(x, P, prefix, training, taps, stats_out=None):
    """nn.BatchNorm2d(eps=1e-3) forward: batch statistics in training mode (the functional form of
    what the reference's modules call), running statistics otherwise."""
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    if training:
        if stats_out is not None:
            stats_out[prefix] = (x.detach().mean(dim=(0, 2, 3)), x.detach().var(dim=(0, 2, 3), unbiased=True))
        return F.batch_norm(x, None, None, w, b, True, 0.1, BN_EPS)
    return F.batch_norm(x, P[prefix + ".running_mean"], P[prefix + ".running_var"], w, b, False, 0.1, BN_EPS)


def downsampler(x, P, prefix, training=True, taps=None, stats_out=None):
    c = F.conv2d(x, P[prefix + ".conv.weight"], P[prefix + ".conv.bias"], stride=2, padding=1)
    p = F.max_pool2d(x, 2, stride=2)
    y = torch.cat([c, p], 1)
    y = F.relu(_bn(y, P, prefix + ".bn", training, taps, stats_out))
    return y


def non_bottleneck_1d(x, P, prefix, dil, training=True, drop_mask=None, taps=None, stats_out=None):
    o = F.conv2d(x, P[prefix + ".conv3x1_1.weight"], P[prefix + ".conv3x1_1.bias"], padding=(1, 0))
    o = F.relu(o)
    o = F.conv2d(o, P[prefix + ".conv1x3_1.weight"], P[prefix + ".conv1x3_1.bias"], padding=(0, 1))
    o = F.relu(_bn(o, P, prefix + ".bn1", training, taps, stats_out))
    o = F.conv2d(o, P[prefix + ".conv3x1_2.weight"], P[prefix + ".conv3x1_2.bias"],
                 padding=(dil, 0), dilation=(dil, 1))
    o = F.relu(o)
    o = F.conv2d(o, P[prefix + ".conv1x3_2.weight"], P[prefix + ".conv1x3_2.bias"],
                 padding=(0, dil), dilation=(1, dil))
    o = _bn(o, P, prefix + ".bn2", training, taps, stats_out)
    if drop_mask is not None:
        o = o * drop_mask[:, :, None, None].to(o.dtype)
    return F.relu(o + x)


def upsampler(x, P, prefix, training=True, taps=None, stats_out=None):
    o = F.conv_transpose2d(x, P[prefix + ".conv.weight"], P[prefix + ".conv.bias"],
                           stride=2, padding=1, output_padding=1)
    return F.relu(_bn(o, P, prefix + ".bn", training, taps, stats_out))


def erfnet_forward(x, P, training=True, drop_masks=None, taps=None, stats_out=None):
    """-> (encoder_output [B,128,H/8,W/8], decoder_output [B,L,H,W]).

    ``taps``: optional dict filled with every block output (for layer-wise parity).
    ``stats_out``: optional dict filled with per-BN (batch mean, unbiased batch var).
    """
    drop_masks = drop_masks or {}

    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t

    o = tap("encoder.initial_block", downsampler(x, P, "encoder.initial_block", training, taps, stats_out))
    o = tap("encoder.layers.0", downsampler(o, P, "encoder.layers.0", training, taps, stats_out))
    for prefix, d in ENC_NB[:5]:
        o = tap(prefix, non_bottleneck_1d(o, P, prefix, d, training, drop_masks.get(prefix), taps, stats_out))
    o = tap("encoder.layers.6", downsampler(o, P, "encoder.layers.6", training, taps, stats_out))
    for prefix, d in ENC_NB[5:]:
        o = tap(prefix, non_bottleneck_1d(o, P, prefix, d, training, drop_masks.get(prefix), taps, stats_out))
    enc = o
    o = tap("decoder.layers.0", upsampler(o, P, "decoder.layers.0", training, taps, stats_out))
    o = tap("decoder.layers.1", non_bottleneck_1d(o, P, "decoder.layers.1", 1, training, None, taps, stats_out))
    o = tap("decoder.layers.2", non_bottleneck_1d(o, P, "decoder.layers.2", 1, training, None, taps, stats_out))
    o = tap("decoder.layers.3", upsampler(o, P, "decoder.layers.3", training, taps, stats_out))
    o = tap("decoder.layers.4", non_bottleneck_1d(o, P, "decoder.layers.4", 1, training, None, taps, stats_out))
    o = tap("decoder.layers.5", non_bottleneck_1d(o, P, "decoder.layers.5", 1, training, None, taps, stats_out))
    dec = F.conv_transpose2d(o, P["decoder.output_conv.weight"], P["decoder.output_conv.bias"], stride=2)
    tap("decoder.output_conv", dec)
    return enc, dec


def full_step(x, P, grid, order, nclasses, zero_rows, x_gt, valid, act="square", const=255.0,
              reg_ls=0.0, resize=256, drop_masks=None, taps=None, stats_out=None, loss_obj=None,
              skip_rows=0):
    """Forward of the whole hot path (Backprojection_Loss/main.py:286-305):
    ERFNet -> activation -> mask -> WLS -> backprojection loss averaged over lanes.
    Returns (loss, beta[B,L,d+1] float64, dec_out, masked).  P must require grad for
    backward; call ``loss.backward()`` on the result."""
    from . import lsq_oracle as lo
    enc, dec = erfnet_forward(x, P, True, drop_masks, taps, stats_out)
    masked = lo.activate_and_mask(dec, act, zero_rows)
    beta, _ = lo.wls_forward(masked, grid, order, const, reg_ls, skip_rows=skip_rows)
    beta = beta.double()                                           # LSQ_layer.py:154
    crit = loss_obj or lo.BackprojectionLoss(order, resize)
    total = 0
    for l in range(nclasses):
        ll, _ = crit(beta[:, l], x_gt[:, l], valid[:, l])
        total = total + ll
    loss = total / nclasses                                        # main.py:305
    return loss, beta, dec, masked
