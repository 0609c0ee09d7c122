"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float32) of LDR-FLIP as the reference computes it: flip-cuda
(/root/reference/ext/flip-cuda: cudaImage.cuh:399-523 filters and driver, cudaKernels.cuh:248-470 passes, color.cuh:205-385 colour
transforms, common/sharedflip.h constants) driven by NeuralBVHRenderer::compute_and_display_flip_error
(/root/reference/src/cuda/neural/neural_bvh_renderer.cu:34-96).  Pinned by execution of that code: tests/golden/ref_flip_*.npz hold
error maps written by oracle/ref_tools/flip_harness.cu (flip-cuda itself) on a B200.  Only tests/ may import this file."""
import numpy as np

F = np.float32
QC, PC, PT, GW, QF = F(0.7), F(0.4), F(0.95), F(0.082), F(0.5)
PI = F(3.14159265358979)
DEFAULT_PPD = float(F(0.7) * (F(3840.0) / F(0.7)) * (PI / F(180.0)))
_M = np.array([[10135552.0 / 24577794.0, 8788810.0 / 24577794.0, 4435075.0 / 24577794.0],
               [2613072.0 / 12288897.0, 8788810.0 / 12288897.0, 887015.0 / 12288897.0],
               [1425312.0 / 73733382.0, 8788810.0 / 73733382.0, 70074185.0 / 73733382.0]], dtype=F)
_MI = np.array([[3.241003275, -1.537398934, -0.498615861], [-0.969224334, 1.875930071, 0.041554224],
                [0.055639423, -0.204011202, 1.057148933]], dtype=F)
_ILL = np.array([0.950428545, 1.0, 1.088900371], dtype=F)
_INV_ILL = np.array([1.052156925, 1.0, 0.918357670], dtype=F)


def linear_to_srgb(l):
    l = l.astype(F)
    with np.errstate(invalid="ignore"):
        return np.where(l <= F(0.0031308), l * F(12.92), F(1.055) * np.power(l, F(1.0 / 2.4), dtype=F) - F(0.055)).astype(F)


def srgb_to_linear(s):
    s = s.astype(F)
    with np.errstate(invalid="ignore"):
        return np.where(s <= F(0.04045), s / F(12.92), np.power((s + F(0.055)) / F(1.055), F(2.4), dtype=F)).astype(F)


def _mat(m, c):
    return (m[:, 0] * c[..., 0:1] + m[:, 1] * c[..., 1:2] + m[:, 2] * c[..., 2:3]).astype(F)


def xyz_to_ycxcz(c):
    c = c * _INV_ILL
    return np.stack([F(116) * c[..., 1] - F(16), F(500) * (c[..., 0] - c[..., 1]), F(200) * (c[..., 1] - c[..., 2])], -1).astype(F)


def ycxcz_to_xyz(c):
    y = (c[..., 0] + F(16)) / F(116)
    return (np.stack([y + c[..., 1] / F(500), y, y - c[..., 2] / F(200)], -1) * _ILL).astype(F)


def xyz_to_lab(c):
    delta = F(6.0 / 29.0)
    d3 = delta * delta * delta
    c = (c * _INV_ILL).astype(F)
    with np.errstate(invalid="ignore"):
        f = np.where(c > d3, np.power(np.maximum(c, F(0)), F(1.0 / 3.0), dtype=F), (F(1.0) / (F(3.0) * delta * delta)) * c + F(4.0 / 29.0)).astype(F)
    return np.stack([F(116) * f[..., 1] - F(16), F(500) * (f[..., 0] - f[..., 1]), F(200) * (f[..., 1] - f[..., 2])], -1).astype(F)


def hunt(lab):
    return np.stack([lab[..., 0], F(0.01) * lab[..., 0] * lab[..., 1], F(0.01) * lab[..., 0] * lab[..., 2]], -1).astype(F)


def hyab(a, b):
    return (np.abs(a[..., 0] - b[..., 0]) + np.sqrt((a[..., 1] - b[..., 1]) ** 2 + (a[..., 2] - b[..., 2]) ** 2)).astype(F)


def filters(ppd):
    ppd = F(ppd)
    rs = int(np.ceil(F(3.0) * np.sqrt(F(0.04) / (F(2.0) * PI * PI)) * ppd))
    x2 = ((np.arange(-rs, rs + 1).astype(F) * (F(1.0) / ppd)) ** 2).astype(F)
    g = lambda a, b: (F(a) * np.sqrt(PI / F(b)) * np.exp(-(PI * PI) * x2 / F(b))).astype(F)
    gs = lambda a, b: (np.sqrt(F(a) * np.sqrt(PI / F(b))) * np.exp(-(PI * PI) * x2 / F(b))).astype(F)
    wy, wcx, wz1, wz2 = g(1.0, 0.0047), g(1.0, 0.0053), gs(34.1, 0.04), gs(13.5, 0.025)
    wy, wcx = wy / wy.sum(dtype=F), wcx / wcx.sum(dtype=F)
    nz = F(1.0) / np.sqrt(wz1.sum(dtype=F) ** 2 + wz2.sum(dtype=F) ** 2)
    sd = F(0.5) * GW * ppd
    rf = int(np.ceil(F(3.0) * sd))
    xx = np.arange(-rf, rf + 1).astype(F)
    gg = np.exp(-(xx * xx) / (F(2.0) * sd * sd)).astype(F)
    dg, ddg = (-xx * gg).astype(F), ((xx * xx / (sd * sd) - F(1.0)) * gg).astype(F)
    dgn = np.where(dg > 0, dg / dg[dg > 0].sum(dtype=F), dg / (-dg[dg <= 0]).sum(dtype=F)).astype(F)
    ddgn = np.where(ddg > 0, ddg / ddg[ddg > 0].sum(dtype=F), ddg / (-ddg[ddg <= 0]).sum(dtype=F)).astype(F)
    return rs, (wy, wcx, wz1 * nz, wz2 * nz), rf, (gg / gg.sum(dtype=F), dgn, ddgn)


def _conv(img, w, axis):
    """Clamp-to-border correlation of a 2-D float32 image with the odd-length filter w along axis, summed from -r to +r."""
    r = len(w) // 2
    pad = [(0, 0), (0, 0)]
    pad[axis] = (r, r)
    p = np.pad(img, pad, mode="edge")
    out = np.zeros_like(img, dtype=F)
    n = img.shape[axis]
    for k in range(len(w)):
        sl = [slice(None), slice(None)]
        sl[axis] = slice(k, k + n)
        out = (out + w[k] * p[tuple(sl)]).astype(F)
    return out


def flip(reference_linear_rgb, test_linear_rgb, ppd=DEFAULT_PPD):
    """(H, W, 3) linear RGB images -> (H, W) LDR-FLIP error map."""
    rs, (wy, wcx, wz1, wz2), rf, (g, dg, ddg) = filters(ppd)
    lab, feat = [], []
    for img in (reference_linear_rgb, test_linear_rgb):
        ycc = xyz_to_ycxcz(_mat(_M, srgb_to_linear(linear_to_srgb(np.asarray(img, F)))))
        fy = _conv(_conv(ycc[..., 0], wy, 1), wy, 0)
        fcx = _conv(_conv(ycc[..., 1], wcx, 1), wcx, 0)
        fcz = _conv(_conv(ycc[..., 2], wz1, 1), wz1, 0) + _conv(_conv(ycc[..., 2], wz2, 1), wz2, 0)
        rgb = np.clip(_mat(_MI, ycxcz_to_xyz(np.stack([fy, fcx, fcz], -1))), F(0), F(1)).astype(F)
        lab.append(hunt(xyz_to_lab(_mat(_M, rgb))))
        yn = (ycc[..., 0] * F(1.0 / 116.0) + F(16.0 / 116.0)).astype(F)
        gx = _conv(yn, g, 1)
        dx, ddx = _conv(_conv(yn, dg, 1), g, 0), _conv(_conv(yn, ddg, 1), g, 0)
        dy, ddy = _conv(gx, dg, 0), _conv(gx, ddg, 0)
        feat.append((np.sqrt(dx * dx + dy * dy).astype(F), np.sqrt(ddx * ddx + ddy * ddy).astype(F)))
    green = hunt(xyz_to_lab(_mat(_M, np.array([[0, 1, 0]], F))))
    blue = hunt(xyz_to_lab(_mat(_M, np.array([[0, 0, 1]], F))))
    cmax = np.power(hyab(green, blue)[0], QC, dtype=F)
    pccmax = PC * cmax
    cd = np.power(hyab(lab[0], lab[1]), QC, dtype=F)
    cd = np.where(cd < pccmax, cd * (PT / pccmax), PT + ((cd - pccmax) / (cmax - pccmax)) * (F(1.0) - PT)).astype(F)
    fd = np.power(F(1.0 / np.sqrt(2.0)) * np.maximum(np.abs(feat[0][0] - feat[1][0]), np.abs(feat[0][1] - feat[1][1])), QF, dtype=F)
    return np.power(cd, F(1.0) - fd, dtype=F)
