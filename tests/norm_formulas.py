"""Pure-torch (CPU, any dtype) transcription of csrc/norm.cu: 7 sums -> coefficients -> affine.
Used by the CPU tests to check the hand-derived first/second-order formulas against autograd of the
oracle, and by the GPU tests as a second opinion on the kernels."""
import torch


def lrelu(v, alpha):
    return torch.where(v > 0, v, v * alpha)


def dlrelu(v, alpha):
    return torch.where(v > 0, torch.ones_like(v), torch.full_like(v, alpha))


def sums7(a, b=None, c=None, flags=0, alpha=0.0):
    n, ch = a.shape[0], a.shape[-1]
    araw = a.reshape(n, -1, ch)
    av = lrelu(araw, alpha) if flags & 1 else araw
    z = torch.zeros(n, ch, dtype=a.dtype)
    bv = b.reshape(n, -1, ch) if b is not None else None
    cv = c.reshape(n, -1, ch) if c is not None else None
    if cv is not None and flags & 4:
        cv = cv * dlrelu(araw, alpha)
    S = [av.sum(1), bv.sum(1) if bv is not None else z, cv.sum(1) if cv is not None else z, (av * av).sum(1),
         (av * bv).sum(1) if bv is not None else z, (av * cv).sum(1) if cv is not None else z,
         (bv * cv).sum(1) if cv is not None else z]
    return torch.stack(S, -1)


def affine(a, b, c, coef, flags=0, alpha=0.0):
    n, ch = a.shape[0], a.shape[-1]
    shp = a.shape
    araw = a.reshape(n, -1, ch)
    av = lrelu(araw, alpha) if flags & 1 else araw
    k = coef[:, None, :, :]
    r = k[..., 0] * av + k[..., 3]
    if b is not None:
        r = r + k[..., 1] * b.reshape(n, -1, ch)
    if c is not None:
        cv = c.reshape(n, -1, ch)
        if flags & 4:
            cv = cv * dlrelu(araw, alpha)
        r = r + k[..., 2] * cv
    if flags & 2:
        r = r * dlrelu(araw, alpha)
    return r.reshape(shp)


def coef(kind, S, p0, p1, N, eps):
    """Returns (coef0, coef1, out0, out1) exactly as norm_coef_kernel."""
    n, ch = S.shape[0], S.shape[1]
    mu = S[..., 0] / N
    var = (S[..., 3] / N - mu * mu).clamp(min=0)
    z = torch.zeros_like(mu)
    st = lambda *xs: torch.stack(xs, -1)
    if kind == 0:
        d = var.sqrt() + eps
        return st(p0 / d, z, z, p1 - mu * p0 / d), None, None, None
    if kind == 1:
        s = var.sqrt(); d = s + eps; g = p0
        m1 = S[..., 1] / N; m2 = S[..., 4] / N - mu * m1
        ka = -g * m2 / (s * d * d)
        return st(ka, g / d + z, z, -g * m1 / d - ka * mu), None, (N * m2 / d).sum(0), S[..., 1].sum(0)
    if kind == 2:
        s = var.sqrt(); d = s + eps; g = p0
        m1 = S[..., 1] / N; mh = S[..., 2] / N
        m2 = S[..., 4] / N - mu * m1; mh2 = S[..., 5] / N - mu * mh
        Ab = S[..., 6] / N - m1 * mh
        isd2 = 1 / (s * d * d)
        cx = -g * Ab * isd2 + g * m2 * mh2 * (d + 2 * s) / (s ** 3 * d ** 3)
        cg = -g * mh2 * isd2; chh = -g * m2 * isd2; kb = -g * mh2 * isd2
        c0 = st(cx, cg, chh, -cx * mu - cg * m1 - chh * mh)
        c1 = st(kb, z, g / d + z, -g * mh / d - kb * mu)
        return c0, c1, (N * (Ab / d - m2 * mh2 * isd2)).sum(0), None
    if kind == 3:
        return None, None, torch.cat([mu, (var + eps).sqrt()], -1), None
    if kind == 4:
        sd = (var + eps).sqrt(); gm, gs = p0[:, :ch], p0[:, ch:]
        ka = gs / N / sd
        return st(ka, z, z, gm / N - ka * mu), None, None, None
    if kind == 5:
        sd = (var + eps).sqrt(); gs = p0[:, ch:]
        mh = S[..., 1] / N; mhx = S[..., 4] / N - mu * mh
        ka = -gs * mhx / N / sd ** 3; kb = gs / N / sd
        return st(ka, kb, z, -kb * mh - ka * mu), None, torch.cat([mh, mhx / sd], -1), None
    if kind == 6:
        r = (var + eps).rsqrt(); sc, bi = p0[:, :ch], p0[:, ch:]
        ka = r * (1 + sc)
        return st(ka, z, z, bi - mu * ka), None, None, None
    if kind == 7:
        r = (var + eps).rsqrt(); sc = p0[:, :ch]
        m1 = S[..., 1] / N; m2 = S[..., 4] / N - mu * m1
        A = r * (1 + sc); ka = -A * r * r * m2
        return st(ka, A, z, -A * m1 - ka * mu), None, torch.cat([N * m2 * r, S[..., 1]], -1), None
    raise ValueError(kind)
