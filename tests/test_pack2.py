"""The 2-bit host panel format (gb_pack2_*) and the chromosome driver (gb_chrom_run_pack2).

CPU part: the HOST packer against a numpy restatement of the layout in include/gauss_b200.h.
GPU part: a panel filled from pack2 rows must give the same integer statistics and the same bits of z / info
as the same SNPs appended as int8 rows; the chromosome driver must equal gb_batch_* for every group count."""
import numpy as np
import pytest

import gauss_b200 as gb
from gauss_b200 import api, synth
from helpers import small_case


def numpy_pack2(pop_sizes, g):
    rb = api.pack2_row_bytes(pop_sizes)
    exp = np.zeros((g.shape[0], rb * 4), np.uint8)
    col = off = 0
    for m in pop_sizes:
        exp[:, col:col + m] = g[:, off:off + m]
        off += m
        col += (m + 127) // 128 * 128
    return (exp[:, 0::4] | (exp[:, 1::4] << 2) | (exp[:, 2::4] << 4) | (exp[:, 3::4] << 6)).astype(np.uint8)


@pytest.mark.parametrize("pop_sizes", [(5, 130, 7), (128,), (1, 1, 1), (61, 103, 40, 25, 2, 330, 97)])
def test_host_packer_layout(pop_sizes):
    ps = np.array(pop_sizes, np.int32)
    rng = np.random.default_rng(len(pop_sizes))
    g = rng.integers(0, 3, (700, int(ps.sum()))).astype(np.int8)      # >= 512 rows: the threaded path
    assert api.pack2_row_bytes(ps) == sum((m + 127) // 128 * 128 for m in pop_sizes) // 4
    np.testing.assert_array_equal(api.pack2_rows_host(ps, g), numpy_pack2(ps, g))
    chars = (g + 48).astype(np.uint8)                                  # ASCII rows, strided source
    wide = np.zeros((700, chars.shape[1] + 9), np.uint8)
    wide[:, :chars.shape[1]] = chars
    np.testing.assert_array_equal(api.pack2_rows_host(ps, wide[:, :chars.shape[1]], is_ascii=True), numpy_pack2(ps, g))
    np.testing.assert_array_equal(api.pack2_rows_host(ps, g[:3]), numpy_pack2(ps, g[:3]))   # single-threaded path
    assert api.pack2_rows_host(ps, g[:0]).shape == (0, api.pack2_row_bytes(ps))


def test_host_packer_refuses_other_dosages():
    ps = np.array([9, 140], np.int32)
    g = np.ones((4, 149), np.int8)
    for bad in (3, 5, -1):
        g2 = g.copy()
        g2[2, 147] = bad
        with pytest.raises(gb.GaussB200Error) as e:
            api.pack2_rows_host(ps, g2)
        assert e.value.status == api.GB_ERR_UNSUPPORTED
    assert api.pack2_row_bytes(np.array([0, 3], np.int32)) == -1


def numpy_unpack5(pop_sizes, rows5):
    boff = [0]
    for m in pop_sizes:
        boff.append(boff[-1] + ((int(m) + 4) // 5 + 3) // 4 * 4)
    out = []
    for k, m in enumerate(pop_sizes):
        blk = rows5[:, boff[k]:boff[k] + (int(m) + 4) // 5].astype(np.int64)
        out.append(np.stack([(blk // 3 ** q) % 3 for q in range(5)], axis=2).reshape(len(rows5), -1)[:, :int(m)])
    return np.concatenate(out, axis=1).astype(np.int8), boff[-1]


@pytest.mark.parametrize("pop_sizes", [(5, 130, 7), (128,), (1, 1, 1), (61, 103, 40, 25, 2, 330, 97)])
def test_ternary_host_packer_layout(pop_sizes):
    ps = np.array(pop_sizes, np.int32)
    rng = np.random.default_rng(len(pop_sizes) + 5)
    g = rng.integers(0, 3, (700, int(ps.sum()))).astype(np.int8)
    rows5 = api.pack5_rows_host(ps, g)
    dec, used = numpy_unpack5(ps, rows5)
    np.testing.assert_array_equal(dec, g)
    assert rows5.shape[1] == api.pack5_row_bytes(ps) == (used + 15) // 16 * 16
    assert (rows5 <= 242).all() and (rows5[:, used:] == 0).all()
    np.testing.assert_array_equal(api.pack5_rows_host(ps, (g + 48).astype(np.uint8), is_ascii=True), rows5)
    np.testing.assert_array_equal(api.pack5_rows_host(ps, g[:3]), rows5[:3])
    for bad in (3, -1):
        g2 = g[:4].copy()
        g2[1, -1] = bad
        with pytest.raises(gb.GaussB200Error):
            api.pack5_rows_host(ps, g2)


@pytest.mark.gpu
def test_pack5_panel_and_chrom_driver(gpu_ctx):
    c = small_case(seed=45, n_snps=900, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 900))
    g, t = c["g"].astype(np.int8), c["type"]
    rows5 = api.pack5_rows_host(c["pop_sizes"], g)
    ref = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "int8")
    ref.append_host(g, is_ascii=False)
    p5 = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
    p5.append_pack5_host(rows5[:333])
    p5.append_pack5_host(rows5[333:])
    a, b = np.arange(0, 600), np.arange(300, 900)
    for x, y in zip(ref.gram_counts(a, b), p5.gram_counts(a, b)):
        np.testing.assert_array_equal(x, y)
    wins = [(0, 300), (150, 520), (400, 900)]
    t_rows, u_rows, t_off, u_off = [], [], [0], [0]
    for lo, hi in wins:
        idx = np.arange(lo, hi)
        t_rows.append(idx[t[lo:hi] == 1])
        u_rows.append(idx[t[lo:hi] == 0])
        t_off.append(t_off[-1] + len(t_rows[-1]))
        u_off.append(u_off[-1] + len(u_rows[-1]))
    rows_t, rows_u = np.concatenate(t_rows), np.concatenate(u_rows)
    batch = gb.Batch(p5, t_off, rows_t, u_off, rows_u, c["z"][rows_t], c["w"])
    batch.run()
    z0, i0, s0 = batch.fetch()
    work = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
    for n_groups in (1, 3):
        z, info, st = work.chrom_run_pack5(rows5.ctypes.data, len(g), rows5.strides[0], t_off, rows_t, u_off, rows_u,
                                           c["z"][rows_t], c["w"], n_groups=n_groups)
        np.testing.assert_array_equal(st, s0)
        np.testing.assert_array_equal(z, z0)
        np.testing.assert_array_equal(info, i0)
    # a byte that is not a code (243..255), and a digit past a population's size, are refused
    for r, col, val in ((10, 0, 250), (20, 12, 242)):       # population 0 has 61 dosages: its 13th byte holds one digit
        bad = rows5.copy()
        bad[r, col] = val
        pb = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
        pb.append_pack5_host(bad)
        _, _, rc = pb.window_distmix(t_rows[0], u_rows[0], c["z"][t_rows[0]], c["w"], allow=(api.GB_ERR_UNSUPPORTED,))
        assert rc == api.GB_ERR_UNSUPPORTED
    with pytest.raises(gb.GaussB200Error):
        ref.append_pack5_host(rows5[:1])


@pytest.mark.gpu
def test_pack2_panel_equals_int8_panel(gpu_ctx):
    c = small_case(seed=41, n_snps=400, pop_sizes=(61, 103, 40, 25, 2, 330, 97))
    g = c["g"].astype(np.int8)
    p_ref = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "int8")
    p_ref.append_host(g, is_ascii=False)
    p2 = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
    rows2 = api.pack2_rows_host(c["pop_sizes"], g)
    p2.append_pack2_host(rows2[:150])
    p2.append_pack2_host(rows2[150:])
    assert p2.n_rows == len(g)
    a, b = np.arange(0, 300), np.arange(100, 400)
    for x, y in zip(p_ref.gram_counts(a, b), p2.gram_counts(a, b)):
        np.testing.assert_array_equal(x, y)
    meas, unme = np.where(c["type"] == 1)[0], np.where(c["type"] == 0)[0]
    p4 = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
    p4.append_host(g, is_ascii=False)
    z_a, i_a, _ = p4.window_distmix(meas, unme, c["z"][meas], c["w"])
    z_b, i_b, _ = p2.window_distmix(meas, unme, c["z"][meas], c["w"])
    np.testing.assert_array_equal(z_a, z_b)
    np.testing.assert_array_equal(i_a, i_b)
    # pack2 rows cannot go into an int8 panel; code 3 is flagged like any dosage E2M1 cannot hold
    with pytest.raises(gb.GaussB200Error):
        p_ref.append_pack2_host(rows2[:1])
    bad = rows2.copy()
    bad[3, 0] |= 3
    p5 = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
    p5.append_pack2_host(bad)
    _, _, rc = p5.window_distmix(meas, unme, c["z"][meas], c["w"], allow=(api.GB_ERR_UNSUPPORTED,))
    assert rc == api.GB_ERR_UNSUPPORTED


@pytest.mark.gpu
@pytest.mark.parametrize("mix", [True, False])
def test_chrom_driver_equals_batch(gpu_ctx, mix):
    c = small_case(seed=43, n_snps=1200, pop_sizes=(61, 103, 40, 25, 2, 330, 97), measured_frac=0.3, core=(0, 1200))
    g, t = c["g"].astype(np.int8), c["type"]
    w = c["w"] if mix else None
    wins = [(0, 300), (150, 520), (400, 900), (880, 900), (700, 1100), (900, 1200)]   # window 3 is too small
    t_rows, u_rows, t_off, u_off = [], [], [0], [0]
    for lo, hi in wins:
        idx = np.arange(lo, hi)
        t_rows.append(idx[t[lo:hi] == 1])
        u_rows.append(idx[t[lo:hi] == 0])
        t_off.append(t_off[-1] + len(t_rows[-1]))
        u_off.append(u_off[-1] + len(u_rows[-1]))
    rows_t, rows_u = np.concatenate(t_rows), np.concatenate(u_rows)
    panel = gb.Panel(gpu_ctx, c["pop_sizes"], len(g), "e2m1")
    panel.append_host(g, is_ascii=False)
    batch = gb.Batch(panel, t_off, rows_t, u_off, rows_u, c["z"][rows_t], w)
    batch.run()
    z0, i0, s0 = batch.fetch()
    assert s0[3] != 0 and (np.delete(s0, 3) == 0).all()
    rows2 = api.pack2_rows_host(c["pop_sizes"], g)
    work = gb.Panel(gpu_ctx, c["pop_sizes"], len(g) + 7, "e2m1")
    for n_groups in (1, 2, 3, 6, 50):
        z, info, st = work.chrom_run_pack2(rows2.ctypes.data, len(g), rows2.strides[0], t_off, rows_t, u_off, rows_u,
                                           c["z"][rows_t], w, n_groups=n_groups)
        np.testing.assert_array_equal(st, s0)
        np.testing.assert_array_equal(z, z0)        # NaN rows of the refused window compare equal too
        np.testing.assert_array_equal(info, i0)
    # rows with a code the format cannot hold are refused for the whole call
    bad = rows2.copy()
    bad[1000, 2] |= 0x0C
    with pytest.raises(gb.GaussB200Error) as e:
        work.chrom_run_pack2(bad.ctypes.data, len(g), bad.strides[0], t_off, rows_t, u_off, rows_u, c["z"][rows_t], w)
    assert e.value.status == api.GB_ERR_UNSUPPORTED
