"""The fused Lanczos step for a diagonal inner-product matrix (kry_lanczos_diag; default for
MINRES / Lanczos with a diagonal ``ip_B`` since round 2, ``KRY_LANCZOS_DIAGB=0`` turns it off) on the
GPU: the kernel against numpy with the same rounding points, and the fixtures with the fused step on
and off."""
import numpy as np
import pytest

import runners

pytestmark = pytest.mark.gpu


def _ctx():
    import torch
    from krypy_b200 import _device
    assert torch.cuda.is_available()
    return _device.Context.get()


def check_lanczos_diag_kernel(ctx, n, dt, with_prev, offset):
    """kry_lanczos_diag against numpy (same rounding points: B q and q rounded to the storage type)"""
    import torch
    tdt = torch.float64 if dt == np.float64 else torch.float32
    rng = np.random.default_rng(n + 7 * offset + with_prev)

    def dev(a):
        # optional misalignment: a view that starts `offset` elements into a larger buffer
        buf = torch.zeros(a.size + 8, dtype=tdt, device=ctx.device)
        v = buf[offset:offset + a.size]
        v.copy_(torch.from_numpy(a).to(ctx.device))
        return v

    vp, vk, q = (rng.standard_normal(n).astype(dt) for _ in range(3))
    b = rng.uniform(1.0, 2.0, n).astype(dt)
    h3 = np.array([0.37 if with_prev else 0.0, 0.25, -1.0])
    vpd, vkd, bd, qd = dev(vp), dev(vk), dev(b), dev(q)
    vnd = dev(np.zeros(n, dtype=dt))
    h3d = torch.from_numpy(h3.copy()).to(ctx.device)
    ctx.lanczos_diag(vpd if with_prev else None, vkd, bd, qd, h3d if with_prev else None, h3d, vnd)
    ctx.sync()
    # numpy restatement
    qq = q.astype(np.float64)
    if with_prev:
        qq = (qq - h3[0] * vp.astype(np.float64)).astype(dt).astype(np.float64)
    bq = (b.astype(np.float64) * qq).astype(dt).astype(np.float64)
    alpha = float(vk.astype(np.float64) @ bq)
    qq = (qq - alpha * vk.astype(np.float64)).astype(dt).astype(np.float64)
    bq = (b.astype(np.float64) * qq).astype(dt).astype(np.float64)
    beta = float(np.sqrt(abs(qq @ bq)))
    rt = 1e-12 if dt == np.float64 else 3e-5
    got = h3d.cpu().numpy()
    scale = np.abs(vk).astype(np.float64) @ np.abs(bq) + 1e-300
    assert got[0] == h3[0]
    assert abs(got[1] - (h3[1] + alpha)) <= rt * scale
    assert abs(got[2] - beta) <= rt * max(beta, 1e-300)
    np.testing.assert_allclose(qd.cpu().numpy(), qq.astype(dt), rtol=rt * 10, atol=rt * 10)
    np.testing.assert_allclose(vnd.cpu().numpy(), (qq / beta).astype(dt), rtol=rt * 10, atol=rt * 10)


@pytest.mark.parametrize("n", [1, 7, 4099, 150001])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
@pytest.mark.parametrize("with_prev", [False, True])
@pytest.mark.parametrize("offset", [0, 1])
def test_lanczos_diag_kernel(n, dt, with_prev, offset):
    check_lanczos_diag_kernel(_ctx(), n, dt, with_prev, offset)


def check_switch_parity(monkeypatch, switch, name, value=True, **kw):
    """a fixture case with a host-level switch set to `value` reproduces the reference"""
    from krypy_b200 import utils
    monkeypatch.setattr(utils, switch, value)
    gold = runners.load_golden(name)
    got = runners.run_product(name, **kw)
    a, b = got["resnorms"], gold["resnorms"]
    assert a.shape == b.shape
    rtol = 1e-5 if name == "shifted_minres_ipB" else 1e-10        # fp32 INPUTS, see test_solvers_gpu.py
    assert np.all(np.abs(a - b) <= rtol * np.abs(b) + 1e-13), float(np.max(np.abs(a - b) / b))
    assert np.abs(got["xk"] - gold["xk"]).max() <= max(1e-8, rtol) * np.abs(gold["xk"]).max()
    return got


@pytest.mark.parametrize("name", ["shifted_minres_ipB_f64", "shifted_minres_ipB"])
def test_fused_diagonal_ipB_lanczos_reproduces_the_reference(monkeypatch, name):
    ctx = _ctx()
    before = ctx.launch_count()
    check_switch_parity(monkeypatch, "_LANCZOS_DIAGB", name, True)
    fused = ctx.launch_count() - before
    before = ctx.launch_count()
    check_switch_parity(monkeypatch, "_LANCZOS_DIAGB", name, False)      # the generic sequence still matches
    assert fused < ctx.launch_count() - before          # one kernel instead of seven per step
