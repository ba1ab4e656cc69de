"""K2 parity on the GPU: exclusion sums and PS against the oracle and the reference-made goldens."""
import numpy as np
import pytest

from oracle import oracle_np
from tests import util

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def _ops():
    from splicedice_b200 import native, ops
    ops.require_cuda()
    return native, ops


def _run(counts, csr, low=None, flags=0, f64=True):
    native, ops = _ops()
    dev = torch.device("cuda", 0)
    J, S = counts.shape
    # leading dimension padded to a multiple of 4 elements: the tiled kernel's 16-byte row alignment
    buf = torch.zeros((J, (S + 3) // 4 * 4), dtype=torch.int32, device=dev)
    buf[:, :S] = torch.from_numpy(counts.astype(np.int32)).to(dev)
    c = buf[:, :S]
    m = None if low is None else torch.from_numpy(low.astype(np.uint8)).to(dev)
    r = ops.quant_ps(c, csr["row_ptr"], csr["col_idx"], low_mask=m, want_f32=True, want_f64=f64, want_exc=True,
                     flags=flags)
    torch.cuda.synchronize()
    return {k: (None if v is None else v.cpu().numpy()) for k, v in r.items()}


@pytest.mark.parametrize("name", ["quant_adversarial.npz", "quant_synth_3k.npz"])
@pytest.mark.parametrize("variant", ["tiled", "gather"])
def test_golden_ps_bits(name, variant):
    """float32 PS bit-for-bit (NaN payloads included) vs SPLICEDICE.calculatePsi run on the reference."""
    native, _ = _ops()
    g = util.load_npz(name)
    counts = g["counts"]
    csr = dict(row_ptr=g["row_ptr"].astype(np.int32), col_idx=g["col_idx"].astype(np.int32))
    flags = native.SD_QUANT_TILED if variant == "tiled" else native.SD_QUANT_GATHER
    got = _run(counts, csr, flags=flags)
    np.testing.assert_array_equal(util.bits32(got["ps_f32"]), g["psi_nolow_bits"])
    low = np.zeros(counts.shape, dtype=np.uint8)
    low[g["low"][:, 0], g["low"][:, 1]] = 1
    got = _run(counts, csr, low=low, flags=flags)
    np.testing.assert_array_equal(util.bits32(got["ps_f32"]), g["psi_bits"])
    np.testing.assert_array_equal(got["exc"], oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"]))


@pytest.mark.parametrize("shape", [(5000, 8), (3000, 64), (2500, 130), (1200, 1000), (700, 1001), (64, 3), (1, 1),
                                   (4000, 260), (900, 100), (500, 65), (33, 193), (2100, 512)])
@pytest.mark.parametrize("variant", ["tiled", "gather"])
def test_oracle_parity(shape, variant):
    native, ops = _ops()
    J, S = shape
    _, csr, counts = util.synthetic_problem(J, S, seed=J + S)
    dev = torch.device("cuda", 0)
    # pad the leading dimension to a multiple of 4 so that the tiled kernel applies to any S
    ld = (S + 3) // 4 * 4 if variant == "tiled" else S
    buf = torch.zeros((J, ld), dtype=torch.int32, device=dev)
    buf[:, :S] = torch.from_numpy(counts).to(dev)
    view = buf[:, :S]
    flags = native.SD_QUANT_TILED if variant == "tiled" else native.SD_QUANT_GATHER
    r = ops.quant_ps(view, csr["row_ptr"], csr["col_idx"], want_f32=True, want_f64=True, want_exc=True, flags=flags)
    torch.cuda.synchronize()
    exc = oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(r["exc"].cpu().numpy(), exc)
    want32 = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"], exc=exc)
    want64 = oracle_np.ps_f64(counts, csr["row_ptr"], csr["col_idx"], exc=exc)
    np.testing.assert_array_equal(util.bits32(r["ps_f32"].cpu().numpy()), util.bits32(want32))
    np.testing.assert_array_equal(util.bits64(r["ps_f64"].cpu().numpy()), util.bits64(want64))


def test_large_counts_take_the_f64_path():
    """Totals above 2^24 must still equal float32(float64 divide)."""
    native, ops = _ops()
    J, S = 600, 40
    _, csr, counts = util.synthetic_problem(J, S, seed=5)
    rng = np.random.default_rng(0)
    counts = (counts.astype(np.int64) * rng.integers(1, 400000, size=counts.shape)).clip(0, 2 ** 31 - 1).astype(np.int32)
    got = _run(counts, csr)
    exc = oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(got["exc"], exc)
    inc = counts.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = (inc / (inc + exc)).astype(np.float32)
    np.testing.assert_array_equal(util.bits32(got["ps_f32"]), util.bits32(want))


def test_row_ranges_and_empty():
    native, ops = _ops()
    J, S = 900, 24
    _, csr, counts = util.synthetic_problem(J, S, seed=77)
    dev = torch.device("cuda", 0)
    c = torch.from_numpy(counts).to(dev)
    out = torch.full((J, S), -5.0, dtype=torch.float32, device=dev)
    ops.quant_ps(c, csr["row_ptr"], csr["col_idx"], out_f32=out, row_begin=100, row_end=333)
    ops.quant_ps(c, csr["row_ptr"], csr["col_idx"], out_f32=out, row_begin=333, row_end=333)   # empty range
    torch.cuda.synchronize()
    want = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"])
    got = out.cpu().numpy()
    np.testing.assert_array_equal(util.bits32(got[100:333]), util.bits32(want[100:333]))
    assert (got[:100] == -5.0).all() and (got[333:] == -5.0).all()
    with pytest.raises(native.NativeCallError):
        ops.quant_ps(c, csr["row_ptr"], csr["col_idx"], row_begin=5, row_end=J + 1)


def test_host_pipeline_matches_device_path():
    native, ops = _ops()
    J, S = 20000, 333
    _, csr, counts = util.synthetic_problem(J, S, seed=3)
    low = (np.random.default_rng(1).random((J, S)) < 0.01).astype(np.uint8)
    pinned = torch.from_numpy(counts).pin_memory()
    out = ops.quant_ps_host(pinned, csr["row_ptr"], csr["col_idx"], low_mask=low).numpy()
    want = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"], low_mask=low)
    np.testing.assert_array_equal(util.bits32(out), util.bits32(want))


@pytest.mark.parametrize("u16", ["1", "0"])
def test_host_pipeline_blocks_ring_and_wide_values(u16, monkeypatch):
    """sd_quant_ps_host over many small row blocks (the pinned staging ring wraps several times),
    with the uint16 link format on and off, on a matrix where some blocks hold counts of 65,536
    and above or exactly 65,535 (those blocks must cross as int32 / still fit), an unaligned
    sample count and a strided host matrix: PS bits equal the oracle's in every case, twice in a
    row (the per-device context and its pool are reused), and after a trim."""
    native, ops = _ops()
    J, S = 30000, 301
    _, csr, counts = util.synthetic_problem(J, S, seed=13)
    counts = counts.astype(np.int32)
    counts[5000:5003, 7] = [65535, 65536, 1 << 20]
    counts[17000, :] = 70000
    counts[29999, 300] = 65535
    want = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"])
    monkeypatch.setenv("SD_QUANT_HOST_BLOCK_MB", "1")
    monkeypatch.setenv("SD_QUANT_HOST_U16", u16)
    pinned = torch.from_numpy(counts).pin_memory()
    for _ in range(2):
        out = ops.quant_ps_host(pinned, csr["row_ptr"], csr["col_idx"]).numpy()
        np.testing.assert_array_equal(util.bits32(out), util.bits32(want))
    native.call("sd_host_pipeline_trim", 0)
    wide = np.zeros((J, S + 11), dtype=np.int32)                 # ld_counts > n_samples, pageable memory
    wide[:, :S] = counts
    out = ops.quant_ps_host(torch.from_numpy(wide)[:, :S], csr["row_ptr"], csr["col_idx"]).numpy()
    np.testing.assert_array_equal(util.bits32(out), util.bits32(want))
    tiny = ops.quant_ps_host(counts[:3].copy(), np.zeros(4, dtype=np.int32), np.zeros(0, dtype=np.int32)).numpy()
    assert np.isnan(tiny[counts[:3] == 0]).all() and (tiny[counts[:3] > 0] == 1.0).all()


def test_device_synth_matches_host_generator():
    native, ops = _ops()
    from splicedice_b200 import synth
    got = ops.synth_counts(seed=9, row0=1000, n_rows=257, n_cols=1000, logical_cols=1000).cpu().numpy()
    want = synth.counts_host(9, 1000, 257, 1000)
    np.testing.assert_array_equal(got, want)


def test_small_divide_exhaustive_and_random():
    """The float32 fast path (div_small) against float32(float64 a / float64 b): every pair
    0 <= a <= b <= 2048 and 4M random pairs with b <= 2^24, through both wide instantiations."""
    native, ops = _ops()
    b = np.repeat(np.arange(1, 2049), np.arange(2, 2050))
    a = np.concatenate([np.arange(0, k + 1) for k in range(1, 2049)])
    rng = np.random.default_rng(4)
    rb = rng.integers(1, 2 ** 24 + 1, size=4_000_000)
    ra = (rng.random(rb.size) * (rb + 1)).astype(np.int64).clip(0, rb)
    edge = np.array([[2 ** 24, 2 ** 24], [2 ** 24 - 1, 2 ** 24], [1, 2 ** 24], [0, 2 ** 24], [2 ** 23 + 1, 2 ** 24 - 1]])
    a = np.concatenate([a, ra, edge[:, 0]]); b = np.concatenate([b, rb, edge[:, 1]])
    S = 1024
    pad = (-a.size) % S
    a = np.concatenate([a, np.zeros(pad, dtype=np.int64)]); b = np.concatenate([b, np.ones(pad, dtype=np.int64)])
    M = a.size // S
    counts = np.empty((2 * M, S), dtype=np.int32)
    counts[0::2] = a.reshape(M, S)
    counts[1::2] = (b - a).reshape(M, S)
    row_ptr = np.arange(2 * M + 1, dtype=np.int32)
    col_idx = (np.arange(2 * M, dtype=np.int32) ^ 1)
    dev = torch.device("cuda", 0)
    c = torch.from_numpy(counts).to(dev)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = (counts.astype(np.float64) / np.repeat(b.reshape(M, S), 2, axis=0)).astype(np.float32)
    lean = ops.quant_ps(c, row_ptr, col_idx)["ps_f32"].cpu().numpy()
    np.testing.assert_array_equal(util.bits32(lean), util.bits32(want))
    general = ops.quant_ps(c, row_ptr, col_idx, want_exc=True)["ps_f32"].cpu().numpy()
    np.testing.assert_array_equal(util.bits32(general), util.bits32(want))


@pytest.mark.parametrize("log_r", [3, 5, 7, -24, -48, -100])
def test_wide_kernel_tile_shapes_and_unstaged_adjacency(log_r):
    """Deep nests (adjacency of one tile larger than the staging buffer) and other tile heights
    (negative: SD_QUANT_ROWS(n), tile heights that are not powers of two -- 48 is what large
    matrices get by default)."""
    native, ops = _ops()
    tile_flag = (log_r << 8) if log_r > 0 else ((-log_r) << 24)
    from splicedice_b200 import synth
    js = [("chr1", 1000, 900000, "+")] + [("chr1", 2000 + 40 * i, 2030 + 40 * i, "+") for i in range(3000)]
    js += synth.junction_tuples(2000, 5)
    c, s, st, en, _, _ = oracle_np.junctions_to_arrays(js)
    csr = oracle_np.cluster_csr(c, s, st, en)
    J, S = len(js), 200
    counts = synth.counts_host(3, 0, J, S)
    dev = torch.device("cuda", 0)
    r = ops.quant_ps(torch.from_numpy(counts).to(dev), csr["row_ptr"], csr["col_idx"], want_exc=True,
                     flags=native.SD_QUANT_TILED | tile_flag)
    exc = oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(r["exc"].cpu().numpy(), exc)
    want = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"], exc=exc)
    np.testing.assert_array_equal(util.bits32(r["ps_f32"].cpu().numpy()), util.bits32(want))
    lean = ops.quant_ps(torch.from_numpy(counts).to(dev), csr["row_ptr"], csr["col_idx"],
                        flags=native.SD_QUANT_TILED | tile_flag)["ps_f32"].cpu().numpy()
    np.testing.assert_array_equal(util.bits32(lean), util.bits32(want))


def test_narrow_tile_kernel_on_wide_matrix():
    native, ops = _ops()
    J, S = 1500, 300
    _, csr, counts = util.synthetic_problem(J, S, seed=8)
    dev = torch.device("cuda", 0)
    r = ops.quant_ps(torch.from_numpy(counts).to(dev), csr["row_ptr"], csr["col_idx"], want_exc=True,
                     flags=native.SD_QUANT_TILED | native.SD_QUANT_NARROW_TILES)
    want = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(util.bits32(r["ps_f32"].cpu().numpy()), util.bits32(want))


def test_sums_beyond_32_bits_take_the_wide_path():
    """Counts near 2^31 with several neighbours: exclusion sums exceed 2^32 and must stay exact."""
    native, ops = _ops()
    J, S = 400, 160
    _, csr, counts = util.synthetic_problem(J, S, seed=12)
    rng = np.random.default_rng(3)
    big = rng.integers(2 ** 30, 2 ** 31 - 1, size=counts.shape)
    counts = np.where(rng.random(counts.shape) < 0.5, big, counts).astype(np.int32)
    got = _run(counts, csr)
    exc = oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"])
    assert exc.max() > 2 ** 32
    np.testing.assert_array_equal(got["exc"], exc)
    inc = counts.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        want64 = inc / (inc + exc)
    np.testing.assert_array_equal(util.bits64(got["ps_f64"]), util.bits64(want64))
    np.testing.assert_array_equal(util.bits32(got["ps_f32"]), util.bits32(want64.astype(np.float32)))


def test_f64_divide_exhaustive_and_random():
    """The float64 lean epilogue (div_fast: reciprocal + residual correction without the range
    check) against numpy's float64 divide: every pair 0 <= a <= b <= 2048 and 4M random pairs
    with b up to 2^32 - 2, through the lean and the general epilogue."""
    native, ops = _ops()
    b = np.repeat(np.arange(1, 2049), np.arange(2, 2050))
    a = np.concatenate([np.arange(0, k + 1) for k in range(1, 2049)])
    rng = np.random.default_rng(6)
    hi = rng.integers(0, 2 ** 31, size=4_000_000)
    lo = rng.integers(0, 2 ** 31, size=4_000_000)
    small = rng.random(hi.size) < 0.5
    hi[small] >>= rng.integers(0, 31, size=int(small.sum()))
    lo[small] >>= rng.integers(0, 31, size=int(small.sum()))
    edge = np.array([[2 ** 31 - 1, 2 ** 31 - 1], [0, 2 ** 31 - 1], [2 ** 31 - 1, 0], [1, 2 ** 31 - 1], [0, 0], [3, 0]])
    inc = np.concatenate([a, hi, edge[:, 0]]); exc = np.concatenate([b - a, lo, edge[:, 1]])
    S = 1024
    pad = (-inc.size) % S
    inc = np.concatenate([inc, np.zeros(pad, dtype=np.int64)]); exc = np.concatenate([exc, np.ones(pad, dtype=np.int64)])
    M = inc.size // S
    counts = np.empty((2 * M, S), dtype=np.int32)
    counts[0::2] = inc.reshape(M, S)
    counts[1::2] = exc.reshape(M, S)
    row_ptr = np.arange(2 * M + 1, dtype=np.int32)
    col_idx = (np.arange(2 * M, dtype=np.int32) ^ 1)
    c = torch.from_numpy(counts).cuda()
    tot = np.repeat((inc + exc).reshape(M, S), 2, axis=0).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        want = counts.astype(np.float64) / tot
    lean = ops.quant_ps(c, row_ptr, col_idx, want_f32=False, want_f64=True)["ps_f64"].cpu().numpy()
    np.testing.assert_array_equal(lean.view(np.uint64), want.view(np.uint64))
    general = ops.quant_ps(c, row_ptr, col_idx, want_f32=False, want_f64=True,
                           flags=native.SD_QUANT_GENERAL)["ps_f64"].cpu().numpy()
    np.testing.assert_array_equal(general.view(np.uint64), want.view(np.uint64))


@pytest.mark.parametrize("shape", [(5000, 200), (1200, 1000), (700, 1001), (33, 193), (2100, 512), (100, 260),
                                   (47, 1000), (49, 300), (1, 257)])
def test_lean_single_output_forms(shape):
    """One output at a time (the lean binary32 / binary64 epilogues) against the oracle bit for bit."""
    native, ops = _ops()
    J, S = shape
    _, csr, counts = util.synthetic_problem(J, S, seed=3 * J + S, zero_frac=0.3)
    dev = torch.device("cuda", 0)
    buf = torch.zeros((J, (S + 3) // 4 * 4), dtype=torch.int32, device=dev)
    buf[:, :S] = torch.from_numpy(counts).to(dev)
    view = buf[:, :S]
    flags = native.SD_QUANT_TILED
    out32 = torch.full((J, (S + 3) // 4 * 4), 7.0, dtype=torch.float32, device=dev)
    out64 = torch.full((J, (S + 3) // 4 * 4), 7.0, dtype=torch.float64, device=dev)
    ops.quant_ps(view, csr["row_ptr"], csr["col_idx"], out_f32=out32[:, :S], flags=flags)
    ops.quant_ps(view, csr["row_ptr"], csr["col_idx"], want_f32=False, out_f64=out64[:, :S], flags=flags)
    exc = oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"])
    want32 = oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"], exc=exc)
    want64 = oracle_np.ps_f64(counts, csr["row_ptr"], csr["col_idx"], exc=exc)
    np.testing.assert_array_equal(util.bits32(out32[:, :S].cpu().numpy()), util.bits32(want32))
    np.testing.assert_array_equal(util.bits64(out64[:, :S].cpu().numpy()), util.bits64(want64))
    assert (out32[:, S:] == 7.0).all() and (out64[:, S:] == 7.0).all()          # padding untouched


def test_lean_forms_on_adversarial_adjacency():
    """Deep nests (degree >= 100: adjacency lists beyond the staged capacity), touching
    intervals and singletons from the reference-made adversarial fixture, wide random counts."""
    native, ops = _ops()
    g = util.load_npz("quant_adversarial.npz")
    csr = dict(row_ptr=g["row_ptr"].astype(np.int32), col_idx=g["col_idx"].astype(np.int32))
    J = len(csr["row_ptr"]) - 1
    S = 300
    counts = np.random.default_rng(12).integers(0, 60, size=(J, S)).astype(np.int32)
    counts[np.random.default_rng(13).random((J, S)) < 0.4] = 0
    c = torch.from_numpy(counts).cuda()
    flags = native.SD_QUANT_TILED
    got32 = ops.quant_ps(c, csr["row_ptr"], csr["col_idx"], flags=flags)["ps_f32"].cpu().numpy()
    got64 = ops.quant_ps(c, csr["row_ptr"], csr["col_idx"], want_f32=False, want_f64=True, flags=flags)["ps_f64"].cpu().numpy()
    exc = oracle_np.exclusion_sums(counts, csr["row_ptr"], csr["col_idx"])
    np.testing.assert_array_equal(util.bits32(got32), util.bits32(oracle_np.ps_f32(counts, csr["row_ptr"], csr["col_idx"], exc=exc)))
    np.testing.assert_array_equal(util.bits64(got64), util.bits64(oracle_np.ps_f64(counts, csr["row_ptr"], csr["col_idx"], exc=exc)))
