"""Pure-Python geometry of the generic conv / wgrad kernels (LfConvArgs / LfWgradArgs in
include/lanefit_b200.h): which taps, phases and weight packings express each reference
layer and its gradients.  No torch device code here, so the CPU test-suite can check every
plan against torch's own convolutions with a small emulator (tests/test_net_plans.py).

Layers (BP/Networks/ERFNet.py): Conv2d 3x3 s2 p1 (:15), Conv2d 3x1 / 1x3 with dilation
(:29-37), ConvTranspose2d 3x3 s2 p1 op1 (:101).
"""


def _phase(Hs, Ws, taps, osy=1, osx=1, oy0=0, ox0=0, isy=1, isx=1):
    return dict(Hs=Hs, Ws=Ws, osy=osy, osx=osx, oy0=oy0, ox0=ox0, isy=isy, isx=isx, taps=taps)


def conv_out_size(Hin, Win, kh, kw, stride, ph, pw, dh, dw):
    Ho = (Hin + 2 * ph - dh * (kh - 1) - 1) // stride + 1
    Wo = (Win + 2 * pw - dw * (kw - 1) - 1) // stride + 1
    return Ho, Wo


def conv_fwd_plan(Hin, Win, kh, kw, stride=1, ph=0, pw=0, dh=1, dw=1):
    """Conv2d forward: out[oy,ox] = sum_t in[oy*s + kh*d - p, ...] W[t].  One phase.
    Tap = (dy, dx, weight slot) with slot = kh_i*kw + kw_i."""
    Ho, Wo = conv_out_size(Hin, Win, kh, kw, stride, ph, pw, dh, dw)
    taps = [(a * dh - ph, b * dw - pw, a * kw + b) for a in range(kh) for b in range(kw)]
    return [_phase(Ho, Wo, taps, isy=stride, isx=stride)], (Ho, Wo)


def conv_dgrad_plan_s1(H, W, kh, kw, ph, pw, dh, dw):
    """Input gradient of a stride-1 'same' Conv2d: din[y,x] = sum_t dout[y + p - kh*d, ...] W[t]^T."""
    taps = [(ph - a * dh, pw - b * dw, a * kw + b) for a in range(kh) for b in range(kw)]
    return [_phase(H, W, taps)], (H, W)


def transposed_gather_plan(Hsmall, Wsmall, Hbig, Wbig, k, p):
    """big[oy,ox] = sum_{kh,kw : (oy+p-kh), (ox+p-kw) even} small[(oy+p-kh)/2, (ox+p-kw)/2] W[kh,kw].
    This is ConvTranspose2d(k, stride 2, padding p) forward (small = layer input) and also the
    input gradient of Conv2d(k, stride 2, padding p) (small = output gradient).  Four phases,
    one per output parity."""
    phases = []
    for py in (0, 1):
        for px in (0, 1):
            taps = [((py + p - a) // 2, (px + p - b) // 2, a * k + b)
                    for a in range(k) for b in range(k)
                    if (py + p - a) % 2 == 0 and (px + p - b) % 2 == 0]
            Hs, Ws = (Hbig - py + 1) // 2, (Wbig - px + 1) // 2
            if taps and Hs > 0 and Ws > 0:
                phases.append(_phase(Hs, Ws, taps, osy=2, osx=2, oy0=py, ox0=px))
    return phases, (Hbig, Wbig)


def convT_dgrad_plan(Hbig, Wbig, Hsmall, Wsmall, k, p):
    """Input gradient of ConvTranspose2d(k, s2, p): dsmall[j,i] = sum_t dbig[2j - p + kh, 2i - p + kw] W[t]
    -- an ordinary stride-2 convolution over the output gradient."""
    taps = [(a - p, b - p, a * k + b) for a in range(k) for b in range(k)]
    return [_phase(Hsmall, Wsmall, taps, isy=2, isx=2)], (Hsmall, Wsmall)


def conv_wgrad_plan(Hin, Win, kh, kw, stride, ph, pw, dh, dw):
    """dW[t][ci][co] = sum_{n,oy,ox} x[n, oy*s + kh*d - p, ..., ci] * dout[n,oy,ox,co].
    P = layer input (gathered), Q = output gradient (dense)."""
    Ho, Wo = conv_out_size(Hin, Win, kh, kw, stride, ph, pw, dh, dw)
    return dict(Hs=Ho, Ws=Wo, psy=stride, psx=stride, qsy=1, qsx=1,
                ptaps=[(a * dh - ph, b * dw - pw) for a in range(kh) for b in range(kw)],
                qtaps=[(0, 0)] * (kh * kw))


def convT_wgrad_plan(Hin, Win, k, p):
    """ConvTranspose2d(k, s2, p): dW[t][ci][co] = sum_{n,j,i} x[n,j,i,ci] * dout[n, 2j - p + kh, 2i - p + kw, co].
    P = layer input (dense), Q = output gradient (gathered)."""
    return dict(Hs=Hin, Ws=Win, psy=1, psx=1, qsy=2, qsx=2,
                ptaps=[(0, 0)] * (k * k),
                qtaps=[(a - p, b - p) for a in range(k) for b in range(k)])


def pad_to(c, m):
    return ((c + m - 1) // m) * m


def cout_pad(co):
    """Column padding of packed weights = the N tile the conv kernel will pick."""
    if co <= 16:
        return 16
    if co <= 32:
        return 32
    if co <= 64:
        return 64
    return pad_to(co, 128)
