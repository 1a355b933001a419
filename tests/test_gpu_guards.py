"""Out-of-bounds and run-to-run checks of our own -- `pytest -m gpu`.

compute-sanitizer is CLOSED on this GPU pool (`gpurun` answers "compute-sanitizer is closed on this pool and stays
closed", recorded in profiles/r2_sanitizer_closed.txt), so the two things memcheck / racecheck would have looked for
are tested directly:

* OUT-OF-BOUNDS WRITES: every output tensor lives inside a larger buffer between two guard bands filled with a
  sentinel; after the kernel the bands must be untouched (ragged shapes, so that tile tails overhang the tensor);
* RACES: every kernel runs several times on the same inputs and must reproduce its output bit for bit (a race
  between the TMA / bulk-copy engine, the tensor core and the epilogue warps shows up as run-to-run differences --
  this is how r2 found the write-after-read hazard in the staged Dice kernels).
"""
import pytest
import torch

from ct_image_segmentation_b200 import _lib, ops
from ct_image_segmentation_b200.ops import ConvGeom

pytestmark = pytest.mark.gpu
DEV = "cuda"
GUARD = 8192          # elements on either side
SENT = 1234.0


class Guarded:
    """(n, d, h, w, c) channels-last tensor (voxel stride ld >= c) between two sentinel bands."""

    def __init__(self, n, sp, c, dtype, ld=None):
        ld = ld or c
        numel = n * sp[0] * sp[1] * sp[2] * ld
        self.buf = torch.full((numel + 2 * GUARD,), SENT, dtype=dtype, device=DEV)
        self.t = self.buf[GUARD:GUARD + numel].view(n, *sp, ld)[..., :c]

    def fill_random(self):
        self.t.copy_(torch.randn(self.t.shape, device=DEV))
        return self

    def intact(self):
        return bool((self.buf[:GUARD] == SENT).all()) and bool((self.buf[-GUARD:] == SENT).all())


def repeat_equal(fn, outs, reps=3):
    fn()
    torch.cuda.synchronize()
    first = [o.clone() for o in outs]
    for _ in range(reps - 1):
        for o in outs:
            o.fill_(0)
        fn()
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(first, outs)), "run-to-run difference"


CONVS = [
    # cin, cout, k, stride, transposed, n, input spatial (ragged: tiles overhang)      kernel family
    (16, 16, 3, 1, False, 2, (9, 40, 24)),     # sliding-window conv, ragged h / d segments
    (32, 32, 3, 1, False, 1, (11, 24, 40)),    # sliding-window conv 32 channels
    (64, 64, 3, 1, False, 2, (6, 6, 10)),      # streaming conv, ragged tiles
    (128, 128, 3, 1, False, 2, (8, 8, 8)),     # split-K cluster conv
    (16, 32, 3, 2, False, 1, (18, 32, 16)),    # stride-2 conv (hi -> lo)
    (64, 16, 3, 2, True, 1, (9, 32, 16)),      # ConvTranspose sliding kernels
    (384, 64, 3, 2, True, 1, (4, 4, 6)),       # ConvTranspose streaming (parity classes)
    (128, 256, 1, 1, False, 2, (5, 6, 7)),     # 1x1x1 conv
]


@pytest.mark.parametrize("cin,cout,k,s,tr,n,sp", CONVS)
def test_conv_kernels_stay_in_bounds_and_reproduce(cin, cout, k, s, tr, n, sp):
    dt = torch.bfloat16
    g = ConvGeom(3, cin, cout, k, s, tr)
    sp_out = g.out_spatial(*sp)
    torch.manual_seed(1)
    x = Guarded(n, sp, cin, dt).fill_random()
    dy = Guarded(n, sp_out, cout, dt).fill_random()
    y, dx = Guarded(n, sp_out, cout, dt), Guarded(n, sp, cin, dt)
    w = torch.randn((cin, cout) + (k,) * 3 if tr else (cout, cin) + (k,) * 3, device=DEV) * 0.05
    wf = ops.pack_weight(g, _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP, w, dt)
    wd = ops.pack_weight(g, _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD, w, dt)
    b = torch.randn(cout, device=DEV)
    gw = torch.full((w.numel() + 2 * GUARD,), SENT, device=DEV)
    gw_view = gw[GUARD:GUARD + w.numel()].view(w.shape)
    repeat_equal(lambda: ops.conv_fprop(g, x.t, wf, b, y.t), [y.t])
    repeat_equal(lambda: ops.conv_fprop_stats(g, x.t, wf, b, y.t), [y.t])
    repeat_equal(lambda: ops.conv_dgrad(g, dy.t, wd, dx.t), [dx.t])
    repeat_equal(lambda: ops.conv_wgrad(g, x.t, dy.t, want_bias=False, out_w=gw_view), [gw_view])
    assert x.intact() and dy.intact() and y.intact() and dx.intact()
    assert bool((gw[:GUARD] == SENT).all()) and bool((gw[-GUARD:] == SENT).all())
    assert torch.isfinite(y.t.float()).all() and torch.isfinite(dx.t.float()).all() and torch.isfinite(gw_view).all()


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("n,sp,c", [(2, (7, 9, 11), 16), (1, (3, 5, 5), 256), (2, (13, 32, 24), 32)])
def test_norm_kernels_stay_in_bounds_and_reproduce(n, sp, c, dt):
    torch.manual_seed(2)
    x, dy = Guarded(n, sp, c, dt).fill_random(), Guarded(n, sp, c, dt).fill_random()
    y, dx = Guarded(n, sp, c, dt), Guarded(n, sp, c, dt)
    alpha = torch.full((1,), 0.2, device=DEV)
    mean, rstd = ops.instnorm_stats(x.t)
    repeat_equal(lambda: ops.instnorm_prelu_fwd(x.t, mean, rstd, alpha, y.t), [y.t])
    da = []
    repeat_equal(lambda: da.append(ops.instnorm_prelu_bwd(x.t, mean, rstd, alpha, dy.t, dx.t)), [dx.t])
    assert all(torch.equal(da[0], d) for d in da[1:])
    assert x.intact() and dy.intact() and y.intact() and dx.intact()


@pytest.mark.parametrize("n,sp", [(2, (5, 7, 9)), (1, (33, 40, 24)), (3, (1, 17, 300))])
def test_dice_and_window_kernels_stay_in_bounds_and_reproduce(n, sp):
    torch.manual_seed(3)
    z = Guarded(n, sp, 10, torch.bfloat16, ld=16).fill_random()     # the production layout: 16-channel rows
    dz = Guarded(n, sp, 10, torch.bfloat16, ld=16)
    lab = torch.randint(0, 10, (n, *sp), device=DEV, dtype=torch.uint8)
    gi, gp = torch.rand(n, 10, device=DEV), torch.rand(n, 10, device=DEV)
    sums = []
    repeat_equal(lambda: sums.append(ops.softmax_dice_metric_sums(z.t, lab)), [])
    assert all(torch.equal(sums[0][0], s[0]) and torch.equal(sums[0][1], s[1]) for s in sums[1:])
    repeat_equal(lambda: ops.softmax_dice_bwd(z.t, lab, gi, gp, dlogits=dz.t), [dz.t])
    assert _lib.load().b200seg_last_launch() == b"softmax_dice_bwd_ring"
    assert z.intact() and dz.intact()
    # sliding-window accumulate + arg-max into a guarded fp32 accumulator
    from ct_image_segmentation_b200.inference import _CudaOps
    D, H, W = sp[0] + 3, sp[1] + 2, sp[2] + 5
    acc = torch.full((D * H * W * 10 + 2 * GUARD,), SENT, device=DEV)
    cnt = torch.full((D * H * W + 2 * GUARD,), SENT, device=DEV)
    a, c = acc[GUARD:-GUARD].view(D, H, W, 10), cnt[GUARD:-GUARD].view(D, H, W)
    a.zero_()
    c.zero_()
    dev_ops = _CudaOps()
    for j in range(n):
        dev_ops.accumulate(z.t[j], None, a, c, 3, 2, 5)
        dev_ops.accumulate(z.t[j], None, a, c, 0, 0, 0)
    c.clamp_(min=1.0)
    lab1, _ = dev_ops.argmax(a, c, False)
    lab2, _ = dev_ops.argmax(a, c, False)
    assert torch.equal(lab1, lab2) and int(lab1.max()) <= 9
    for g_ in (acc, cnt):
        assert bool((g_[:GUARD] == SENT).all()) and bool((g_[-GUARD:] == SENT).all())
