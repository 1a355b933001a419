"""Index logic of the line-tiled tcgen05 kernel (``csrc/tc_line.cu``) as a numpy model, checked against torch's
convolution on CPU: weight-slot order (kh, output slab, kw), mirrored taps for dgrad, the accumulator-chunk ring with
its wrap split, voxel pairs as MMA rows with the parity selected by the K offset, and the epilogue's neighbour sums
``out[w] = acc_kw0[w-1] + acc_kw1[w] + acc_kw2[w+1]`` with zero padding at the ends of a line.  (What the model cannot
show -- descriptors, swizzles, barriers -- is covered by the ``-m gpu`` parity tests with ``B200SEG_LINE_CONV=1``,
``tests/test_gpu_experimental.py``.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

LACCR, LCHUNK = 5, 48   # csrc/tc_line.cu: output slabs resident in TMEM, columns per slab and voxel parity


def line_conv_model(x, w, flip, dseg):
    """x (N, D, H, W, 16) channels-last, w (cout, cin, 3, 3, 3) torch layout; flip = dgrad (mirrored taps)."""
    N, D, H, W, C = x.shape
    PPL = W // 2
    LPT = 128 // PPL
    wt = np.stack([(w[:, :, kd, kh, kw].T if flip else w[:, :, kd, kh, kw])
                   for kd in range(3) for kh in range(3) for kw in range(3)])            # packed [tap][dst][src]
    slot = np.zeros((27, 16, 16), np.float32)
    for kh in range(3):
        for i in range(3):
            for kw in range(3):
                kd = 2 - i
                tap = ((2 - kd) * 3 + (2 - kh)) * 3 + (2 - kw) if flip else (kd * 3 + kh) * 3 + kw
                slot[(kh * 3 + i) * 3 + kw] = wt[tap]
    tilesH = (H + LPT - 1) // LPT
    nseg = (D + dseg - 1) // dseg
    xp = np.zeros((N, D + 2, H + 2 + LPT, W, C), np.float32)                              # TMA out-of-bounds fill
    xp[:, 1:D + 1, 1:H + 1] = x
    out = np.zeros((N, D, H, W, C), np.float32)
    acc = np.zeros((2, LACCR * LCHUNK, 128), np.float32)                                  # [parity][column][row]
    ob = 0
    rows = np.arange(128)
    pp, ln = rows % PPL, rows // PPL
    for item in range(N * nseg * tilesH):
        n, rem = divmod(item, nseg * tilesH)
        seg, th = divmod(rem, tilesH)
        h0, d_begin = th * LPT, seg * dseg
        nd = min(dseg, D - d_begin)
        for s in range(nd + 2):
            lo, hi = max(s - 2, 0), min(s, nd - 1)
            slab = xp[n, d_begin + s, h0: h0 + LPT + 2]                                    # lines h0 - 1 ...
            cnt = hi - lo + 1
            c_lo = (ob + lo) % LACCR
            len0 = min(cnt, LACCR - c_lo)
            len1 = cnt - len0
            for kh in range(3):
                for r in range(2):
                    a = slab[kh:kh + LPT].reshape(LPT * W, C)[r::2]                        # 128 pair rows, voxel r
                    b0 = kh * 9 + (lo - (s - 2)) * 3
                    acc[r, c_lo * 48:(c_lo + len0) * 48] += (a @ slot[b0:b0 + len0 * 3].reshape(len0 * 48, C).T).T
                    if len1:
                        acc[r, :len1 * 48] += (a @ slot[b0 + len0 * 3:b0 + cnt * 3].reshape(len1 * 48, C).T).T
            if s >= 2:
                j = s - 2
                ch = (ob + j) % LACCR
                e, o = acc[0, ch * 48:(ch + 1) * 48].copy(), acc[1, ch * 48:(ch + 1) * 48].copy()
                acc[:, ch * 48:(ch + 1) * 48] = 0                                          # handed back zeroed
                up = np.roll(o[0:16], 1, axis=1)
                up[:, pp == 0] = 0
                dn = np.roll(e[32:48], -1, axis=1)
                dn[:, pp == PPL - 1] = 0
                f0, f1 = up + e[16:32] + o[32:48], e[0:16] + o[16:32] + dn
                ok = (h0 + ln) < H
                out[n, d_begin + j, (h0 + ln)[ok], 2 * pp[ok]] = f0[:, ok].T
                out[n, d_begin + j, (h0 + ln)[ok], 2 * pp[ok] + 1] = f1[:, ok].T
        ob += nd
    return out


@pytest.mark.parametrize("flip", [False, True])
@pytest.mark.parametrize("n,d,h,w,dseg", [(1, 7, 5, 32, 4), (2, 6, 9, 64, 6), (1, 11, 3, 128, 4)])
def test_line_kernel_decomposition(n, d, h, w, dseg, flip):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((n, d, h, w, 16)).astype(np.float32)
    wgt = (rng.standard_normal((16, 16, 3, 3, 3)) * 0.1).astype(np.float32)
    xt = torch.from_numpy(x).permute(0, 4, 1, 2, 3)
    ref = (F.conv_transpose3d if flip else F.conv3d)(xt, torch.from_numpy(wgt), padding=1).permute(0, 2, 3, 4, 1).numpy()
    out = line_conv_model(x, wgt, flip, dseg)
    assert np.abs(out - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
