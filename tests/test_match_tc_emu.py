"""CPU tier: the tensor-core matcher - the kernel body of csrc/match_tc_kernels.cuh, both instantiations (the
reference's arg-min matcher = match_tc_kernel, and match_tc_top2_kernel of the opt-in pano_match_knn), compiled
UNCHANGED by g++ - on the CPU emulation of the CUDA execution model (tests/hostsim/cuda_emu.hpp) plus a host MODEL of
the Blackwell machinery it drives through inline PTX (tests/hostsim/tcgen05_emu.hpp: mbarrier phases with expect-tx /
complete-tx, TMA 2-D tile loads with 128-byte swizzle, bulk copies, tcgen05.mma kind::i8 from K-major swizzled
descriptors into tensor memory, tcgen05.commit, tcgen05.ld, named barriers).  What this establishes is the kernel's own
logic: super-tile scheduling and runs that cross super-rows, pipeline stages and barrier phases (a wrong wait or release
shows up as a reported deadlock), descriptor arithmetic, the epilogue's packed keys and reductions, padding columns and
rows, the cross-CTA merge (atomicMin / the two-slot publication).  That the hardware agrees with the model is what the
GPU tier establishes (tests/test_gpu_parity.py: tensor-core matcher = SIMT matcher = oracle).
(Run with the emulation's default thread order.  Under PANO_EMU_THREAD_ORDER=1|2 - adversarial scheduling - a consumer
warp can be held back until its stage's barrier has advanced two phases, which the emulation reports as a deadlock:
DESIGN section 9.)"""
import ctypes as C

import numpy as np
import pytest


def p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def descriptors(rng, n, kind="random"):
    d = np.zeros((n, 128), np.uint8)
    if kind == "random":
        d[:, :75] = rng.integers(0, 256, (n, 75))
    elif kind == "quantised":                      # few grey levels: many exact distance ties
        d[:, :75] = rng.integers(0, 3, (n, 75)) * 100
    elif kind == "extreme":
        d[:, :75] = rng.choice([0, 255], (n, 75))
    return d


def brute(qd, td):
    D = ((qd[:, None, :].astype(np.int64) - td[None, :, :].astype(np.int64)) ** 2).sum(-1)
    order = np.lexsort((np.broadcast_to(np.arange(td.shape[0]), D.shape), D), axis=1)
    return D, order


def run(lib, qd, td, ctas, top2, order=2):
    nq, nt = len(qd), len(td)
    b1 = np.zeros(nq, np.uint64)
    b2 = np.zeros(nq, np.uint64)
    st = lib.tcemu_match(p(qd, C.c_uint8), nq, p(td, C.c_uint8), nt, ctas, top2, order, p(b1, C.c_uint64), p(b2, C.c_uint64))
    assert st == 0, (st, lib.tcemu_last_error())
    return b1, b2


def check(lib, qd, td, ctas, top2, order=2):
    b1, b2 = run(lib, qd, td, ctas, top2, order)
    D, o = brute(qd, td)
    for q in range(len(qd)):
        assert int(b1[q]) == (int(D[q, o[q, 0]]) << 32 | int(o[q, 0])), ("nearest", q)
        if top2:
            want = (int(D[q, o[q, 1]]) << 32 | int(o[q, 1])) if len(td) >= 2 else 2 ** 64 - 1
            assert int(b2[q]) == want, ("runner-up", q)


@pytest.mark.parametrize("top2", [0, 1])
@pytest.mark.parametrize("nq,nt,ctas", [(100, 200, 148), (600, 700, 148), (600, 700, 3), (600, 700, 1), (513, 129, 5),
                                        (1, 1, 148), (1, 2, 148), (127, 128, 2), (128, 127, 148), (129, 257, 7),
                                        (1030, 130, 2), (40, 1500, 148)])
def test_emulated_tensor_core_matcher_equals_brute_force(match_tc_emu, nq, nt, ctas, top2):
    rng = np.random.default_rng(nq * 7919 + nt * 13 + ctas)
    qd, td = descriptors(rng, nq), descriptors(rng, nt)
    if nt > 5:
        td[nt // 2] = td[1]                     # an exact duplicate: equal SSD, the earlier train row ranks first
        qd[0] = td[1]                           # SSD 0 twice for query 0
    check(match_tc_emu, qd, td, ctas, top2)


@pytest.mark.parametrize("top2", [0, 1])
@pytest.mark.parametrize("kind", ["quantised", "extreme"])
def test_emulated_tensor_core_matcher_ties_and_extreme_norms(match_tc_emu, kind, top2):
    rng = np.random.default_rng(5)
    qd, td = descriptors(rng, 300, kind), descriptors(rng, 390, kind)
    td[-1, :75] = 255
    qd[-1, :75] = 0                             # the largest SSD the key arithmetic has to hold: 75 * 255^2
    for ctas, order in ((148, 0), (4, 1), (2, 2)):
        check(match_tc_emu, qd, td, ctas, top2, order)


def test_emulated_tensor_core_variants_agree_on_the_nearest_neighbour(match_tc_emu):
    rng = np.random.default_rng(9)
    qd, td = descriptors(rng, 700), descriptors(rng, 520)
    a, _ = run(match_tc_emu, qd, td, 148, 0)
    b, _ = run(match_tc_emu, qd, td, 148, 1)
    assert np.array_equal(a, b)
