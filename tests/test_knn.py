"""CPU tier of the opt-in 2-NN / Lowe-ratio matcher (pano_match_knn; north star item (c); NOT a reference function).

1. The checker (oracle.match_knn) is pinned to the published algorithm it restates: real cv2.BFMatcher.knnMatch(k=2)
   with NORM_L2SQR on the patch bytes and NORM_HAMMING on the binary descriptors, plus Lowe's ratio test, and - with a
   ratio of 1 and two or more candidates - to the reference matcher's own nearest neighbour (oracle.match).
2. The engine's kernels (csrc/knn_kernels.cuh) are compiled UNCHANGED by g++ on a CPU emulation of the CUDA execution
   model (tests/hostsim/cuda_emu.hpp: threads as fibers, real barriers and warp shuffles) and must reproduce the checker
   bit for bit, for every split of the train range and every block order (the two-slot atomic publication).
3. The arithmetic of the tensor-core matcher's top-2 epilogue (tile keys, chains, fold, publication) replayed on the
   host gives the same (nearest, runner-up) keys as the brute force.
The GPU tier (tests/test_zz4_knn_gpu.py) runs the same comparisons through the C ABI."""
import ctypes as C
import itertools

import numpy as np
import pytest

from conftest import load_synth

MATCH_DTYPE = np.dtype([("queryIdx", "<i4"), ("trainIdx", "<i4"), ("distance", "<f4")])


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


@pytest.fixture(scope="module")
def scene(oracle):
    left, right, _ = load_synth().make_pair(480, 300, seed=5)
    kl, kr = oracle.detect(left), oracle.detect(right)
    # a few keypoints that fail the in-border test, and duplicates (exact distance ties)
    kl = np.concatenate([kl, [[0, 0], [1, 299], [479, 5]], kl[:7]]).astype(np.int32)
    kr = np.concatenate([[[2, 1]], kr, kr[3:5]]).astype(np.int32)
    return left, right, kl, kr


def in_border(k, w, h, b=2):
    return ~((k[:, 0] < b) | (k[:, 1] < b) | (k[:, 0] + b >= w) | (k[:, 1] + b >= h))


def patches(img, k, b=2):
    return np.stack([img[y - b:y + b + 1, x - b:x + b + 1].reshape(-1) for x, y in k]) if len(k) else np.zeros((0, 75), np.uint8)


def binary_descriptor_numpy(img, x, y):
    """independent restatement: gray = cvtColor's 15-bit formula, bit k compares pair number 37 k mod 300"""
    pairs = list(itertools.combinations(range(25), 2))
    pt = img[y - 2:y + 3, x - 2:x + 3].astype(np.int64).reshape(25, 3)
    g = (pt[:, 0] * 3735 + pt[:, 1] * 19235 + pt[:, 2] * 9798 + 16384) >> 15
    bits = np.zeros(8, np.uint32)
    for k in range(256):
        a, b = pairs[(37 * k) % 300]
        if g[a] < g[b]:
            bits[k >> 5] |= np.uint32(1 << (k & 31))
    return bits


def brute_top2(D):
    """(d1, j1, d2, j2) per row of a distance matrix, ordered by (distance, column)"""
    out = []
    for row in D:
        order = np.lexsort((np.arange(len(row)), row))
        out.append((int(row[order[0]]), int(order[0]), int(row[order[1]]), int(order[1])))
    return out


# ---- 1. the checker against cv2 -----------------------------------------------------------------------------------
def test_binary_descriptor_matches_its_definition_and_cv2_gray(oracle, scene):
    cv2 = pytest.importorskip("cv2")
    left, _, kl, _ = scene
    gray = cv2.cvtColor(left, cv2.COLOR_BGR2GRAY)
    pairs = list(itertools.combinations(range(25), 2))
    for x, y in kl[in_border(kl, 480, 300)][:60]:
        bits = oracle.knn_binary_descriptor(left, x, y)
        assert np.array_equal(bits, binary_descriptor_numpy(left, x, y))
        g = gray[y - 2:y + 3, x - 2:x + 3].reshape(-1).astype(int)     # real cvtColor
        for k in (0, 1, 37, 128, 255):
            a, b = pairs[(37 * k) % 300]
            assert ((int(bits[k >> 5]) >> (k & 31)) & 1) == int(g[a] < g[b])


@pytest.mark.parametrize("descriptor,ratio", [(0, 0.75), (0, 0.9), (1, 0.8), (1, 0.95)])
def test_oracle_knn_equals_cv2_bfmatcher_with_lowe_ratio(oracle, scene, descriptor, ratio):
    cv2 = pytest.importorskip("cv2")
    left, right, kl, kr = scene
    qi, ti = np.flatnonzero(in_border(kr, 480, 300)), np.flatnonzero(in_border(kl, 480, 300))
    if descriptor == 0:
        dq, dt = patches(right, kr[qi]).astype(np.float32), patches(left, kl[ti]).astype(np.float32)
        bf, factor = cv2.BFMatcher(cv2.NORM_L2SQR), ratio * ratio
    else:
        dq = np.stack([oracle.knn_binary_descriptor(right, x, y) for x, y in kr[qi]]).view(np.uint8)
        dt = np.stack([oracle.knn_binary_descriptor(left, x, y) for x, y in kl[ti]]).view(np.uint8)
        bf, factor = cv2.BFMatcher(cv2.NORM_HAMMING), ratio
    knn = bf.knnMatch(dq, dt, k=2)
    expect = {}
    for a, (m1, m2) in enumerate(knn):
        d1, d2 = int(m1.distance), int(m2.distance)
        assert m1.distance == d1 and m2.distance == d2        # cv2's float distances are exact integers here
        if float(d1) < factor * float(d2):
            expect[int(qi[a])] = (d1, d2, int(ti[m1.trainIdx]))
    m, second = oracle.match_knn(kr, kl, right, left, descriptor=descriptor, ratio=ratio)
    assert len(m) == len(expect) > 5
    assert list(m["queryIdx"]) == sorted(expect)              # ascending query order
    # distances equal cv2's; the neighbour too wherever cv2's choice is not a tie (cv2 leaves tie order unspecified)
    if descriptor == 0:
        D = ((dq[:, None, :] - dt[None, :, :]) ** 2).sum(-1)
    else:
        D = np.unpackbits(dq[:, None, :] ^ dt[None, :, :], axis=-1).sum(-1)
    row_of = {int(q): a for a, q in enumerate(qi)}
    for rec, s in zip(m, second):
        d1, d2, j = expect[int(rec["queryIdx"])]
        assert (int(rec["distance"]), int(s)) == (d1, d2)
        row = D[row_of[int(rec["queryIdx"])]]
        first = int(ti[np.flatnonzero(row == d1)[0]])          # the earliest train keypoint at the nearest distance
        assert int(rec["trainIdx"]) == first
        if (row == d1).sum() == 1:
            assert j == first


def test_oracle_knn_nearest_is_the_reference_matchers_nearest(oracle, scene):
    """ratio = 1 keeps every query whose nearest is strictly nearer than its runner-up: those rows must carry exactly
    the reference matcher's (queryIdx, trainIdx, distance) (ref: src/serial/main.cpp:188-244 via oracle.match)"""
    left, right, kl, kr = scene
    ref = {int(r["queryIdx"]): r for r in oracle.match(kr, kl, right, left)}
    m, second = oracle.match_knn(kr, kl, right, left, descriptor=0, ratio=1.0)
    assert 0 < len(m) <= len(ref)
    for rec, s in zip(m, second):
        r = ref[int(rec["queryIdx"])]
        assert (rec["trainIdx"], rec["distance"]) == (r["trainIdx"], r["distance"]) and rec["distance"] < s
    # the dropped rows are exactly the exact ties (duplicated keypoints in the scene)
    dropped = set(ref) - set(int(q) for q in m["queryIdx"])
    assert len(dropped) >= 1


# ---- 2. the engine's kernels on the CPU emulation of the CUDA execution model ---------------------------------------
def run_emu(lib, kq, kt, imq, imt, patch=5, descriptor=0, ratio=0.75, splits=0, order=0):
    kq = np.ascontiguousarray(kq, np.int32).reshape(-1, 2)
    kt = np.ascontiguousarray(kt, np.int32).reshape(-1, 2)
    imq, imt = np.ascontiguousarray(imq), np.ascontiguousarray(imt)
    out = np.zeros(max(len(kq), 1), MATCH_DTYPE)
    second = np.zeros(max(len(kq), 1), np.float32)
    n = lib.emu_match_knn(p(kq, C.c_int32), len(kq), p(kt, C.c_int32), len(kt),
                          p(imq, C.c_uint8), imq.shape[1], imq.shape[0], C.c_size_t(imq.strides[0]),
                          p(imt, C.c_uint8), imt.shape[1], imt.shape[0], C.c_size_t(imt.strides[0]),
                          patch, descriptor, C.c_double(ratio), splits, order, out.ctypes.data_as(C.c_void_p),
                          p(second, C.c_float), len(out))
    assert n >= 0, (n, lib.emu_last_error())
    return out[:n], second[:n]


@pytest.mark.parametrize("descriptor", [0, 1])
@pytest.mark.parametrize("splits,order", [(0, 0), (1, 0), (3, 1), (7, 2)])
def test_emulated_kernels_equal_the_checker(knn_emu, oracle, scene, descriptor, splits, order):
    left, right, kl, kr = scene
    for ratio in (0.75, 1.0):
        m, s = run_emu(knn_emu, kr, kl, right, left, descriptor=descriptor, ratio=ratio, splits=splits, order=order)
        mo, so = oracle.match_knn(kr, kl, right, left, descriptor=descriptor, ratio=ratio)
        assert len(mo) > 5
        assert np.array_equal(m, mo) and np.array_equal(s, so)


@pytest.mark.parametrize("descriptor", [0, 1])
def test_emulated_kernels_edge_cases(knn_emu, oracle, scene, descriptor):
    left, right, kl, kr = scene
    ok_l = kl[in_border(kl, 480, 300)]
    cases = [(kr, ok_l[:1]),             # one candidate only: no runner-up, no match
             (kr, ok_l[:2]),             # exactly two
             (kr[:1], kl),               # a single query (border keypoint: not a candidate)
             (kr[1:2], kl),              # a single in-border query
             (kr, kl[-10:]),             # train list of duplicates and border points
             (np.zeros((0, 2), np.int32), kl), (kr, np.zeros((0, 2), np.int32)),
             (kr[:130], ok_l[:65]), (kr[:129], ok_l[:64]), (kr[:33], ok_l[:33])]   # ragged tiles / warps
    for kq, kt in cases:
        m, s = run_emu(knn_emu, kq, kt, right, left, descriptor=descriptor, ratio=0.8)
        mo, so = oracle.match_knn(kq, kt, right, left, descriptor=descriptor, ratio=0.8)
        assert np.array_equal(m, mo) and np.array_equal(s, so)
    # flat images: every distance is 0, every query ties on all candidates -> Lowe's strict test rejects everything
    flat = np.full((60, 80, 3), 77, np.uint8)
    k = np.array([[x, y] for y in range(5, 50, 9) for x in range(5, 70, 7)], np.int32)
    m, _ = run_emu(knn_emu, k, k, flat, flat, descriptor=descriptor, ratio=1.0)
    mo, _ = oracle.match_knn(k, k, flat, flat, descriptor=descriptor, ratio=1.0)
    assert len(m) == len(mo) == 0


@pytest.mark.parametrize("patch", [1, 3])
def test_emulated_ssd_kernel_other_patch_sizes(knn_emu, oracle, scene, patch):
    left, right, kl, kr = scene
    m, s = run_emu(knn_emu, kr, kl, right, left, patch=patch, ratio=0.9)
    mo, so = oracle.match_knn(kr, kl, right, left, patch=patch, ratio=0.9)
    assert len(mo) > 0 and np.array_equal(m, mo) and np.array_equal(s, so)


def test_emulated_binary_descriptor_kernel(knn_emu, oracle, scene):
    left, _, kl, _ = scene
    kl = np.ascontiguousarray(kl, np.int32)
    bits = np.zeros((len(kl), 8), np.uint32)
    orig = np.zeros(len(kl), np.int32)
    n = knn_emu.emu_binary_descriptors(p(kl, C.c_int32), len(kl), p(left, C.c_uint8), 480, 300, C.c_size_t(left.strides[0]),
                                       p(bits, C.c_uint32), p(orig, C.c_int32))
    assert n == int(in_border(kl, 480, 300).sum()) > 50
    assert np.array_equal(orig[:n], np.flatnonzero(in_border(kl, 480, 300)))
    for row, i in zip(bits[:n], orig[:n]):
        assert np.array_equal(row, oracle.knn_binary_descriptor(left, kl[i, 0], kl[i, 1]))


def test_emulation_selftests(knn_emu):
    """the emulation itself: barriers and shuffles really exchange data between fibers, and a barrier that cannot
    complete (divergent participants) is reported instead of hanging"""
    rng = np.random.default_rng(1)
    v = rng.integers(-1000, 1000, 5000).astype(np.int32)
    for blocks, threads in ((1, 32), (3, 256), (7, 96), (2, 1024)):
        assert knn_emu.emu_selftest_reduce(p(v, C.c_int32), len(v), blocks, threads) == int(v.sum())
    out = np.zeros(64, np.int32)
    assert knn_emu.emu_selftest_half_warps(p(out, C.c_int32)) == 0          # disjoint groups of one warp meet independently
    assert list(out[:16]) == [120 + 0xffff] * 16 and list(out[16:32]) == [31 + 0xffff] * 16 and list(out[32:]) == list(out[:32])
    assert knn_emu.emu_selftest_deadlock() == 1
    assert b"deadlock" in knn_emu.emu_last_error()


# ---- 3. the tensor-core matcher's top-2 epilogue arithmetic ---------------------------------------------------------
@pytest.mark.parametrize("nq,nt,runs,order", [(40, 300, 1, 0), (40, 300, 3, 1), (25, 129, 2, 2), (10, 128, 1, 0),
                                              (10, 1, 1, 0), (10, 2, 1, 0), (30, 700, 6, 2)])
def test_tc_top2_epilogue_arithmetic_equals_brute_force(knn_emu, nq, nt, runs, order):
    rng = np.random.default_rng(nq * 1000 + nt)
    qd = np.zeros((nq, 128), np.uint8)
    td = np.zeros((nt, 128), np.uint8)
    qd[:, :75] = rng.integers(0, 256, (nq, 75))
    td[:, :75] = rng.integers(0, 256, (nt, 75))
    if nt > 5:
        td[nt // 2] = td[1]                 # an exact duplicate: equal SSD, the earlier column must rank first
        qd[0, :75] = td[1, :75]             # SSD 0 twice for query 0
    td[-1, :75] = 255                       # extreme norms
    qd[-1, :75] = 0
    b1 = np.zeros(nq, np.uint64)
    b2 = np.zeros(nq, np.uint64)
    knn_emu.emu_tc_top2(p(qd, C.c_uint8), nq, p(td, C.c_uint8), nt, runs, order, p(b1, C.c_uint64), p(b2, C.c_uint64))
    D = ((qd[:, None, :].astype(np.int64) - td[None, :, :].astype(np.int64)) ** 2).sum(-1)
    for q in range(nq):
        order_ = np.lexsort((np.arange(nt), D[q]))
        assert int(b1[q]) == (int(D[q, order_[0]]) << 32 | int(order_[0]))
        if nt >= 2:
            assert int(b2[q]) == (int(D[q, order_[1]]) << 32 | int(order_[1]))
        else:
            assert int(b2[q]) == 2 ** 64 - 1
