"""Host-side checks of the stream-K work split used by the experimental prefill kernel (b2q_gemm2s.cu).

`b2q_debug_gemm_plan` runs the SAME SkIter / sk_last_contributor code the kernel's four warp roles run
(gptqmodel_b200/csrc/b2q_streamk.h), so these invariants are checked on the real decomposition:
  * every (tile, k-block) is processed exactly once,
  * a split tile has exactly one owner (the pair holding k-block 0), the owner waits for exactly the number of partial
    accumulators that other pairs park for that tile, and all of them come from pairs AFTER the owner,
  * a pair parks at most one partial (one workspace slot per pair) and does so in its FIRST item; the piece it owns is
    the LAST item of its stream-K segment (so contributors never wait and owners wait last: no cyclic waits),
  * the work is balanced: no pair carries more than one k-block above the average,
  * flags / slots fit the workspace the library asks for.
"""
import ctypes

import pytest

from gptqmodel_b200 import _lib as g

CASES = [(2048, 4096, 4096), (2048, 4096, 1024), (2048, 4096, 14336), (2048, 14336, 4096), (2048, 4096, 6144),
         (2048, 4096, 28672), (129, 4096, 4096), (300, 4096, 4096), (256, 128, 256), (257, 64, 32), (8192, 8192, 8192),
         (4096, 8192, 1280), (1000, 2048, 2016), (2048, 4096, 18944), (512, 11008, 4096), (2048, 4096, 256 * 74),
         (2048, 4096, 256 * 75), (256 * 3, 4096, 256 * 49)]


def items_of(M, K, N, pair):
    plan = (ctypes.c_int * 5)()
    buf = (ctypes.c_int * (4 * 512))()
    n = g.lib.b2q_debug_gemm_plan(M, K, N, pair, plan, buf, 512)
    assert n >= 0
    return list(plan), [tuple(buf[4 * i:4 * i + 4]) for i in range(n)]


@pytest.mark.parametrize("M,K,N", CASES)
def test_streamk_plan_invariants(M, K, N):
    plan, _ = items_of(M, K, N, 0)
    tiles, P, nkb, dp_tiles, sk_tiles = plan
    TM, TN = (M + 255) // 256, (N + 255) // 256
    assert tiles == TM * TN and nkb == K // 64 and dp_tiles + sk_tiles == tiles and 1 <= P <= 74
    assert sk_tiles * 2 * 4 <= 4096                                   # flags region (GEMM2S_FLAG_BYTES)
    assert 4096 + P * 2 * 128 * 256 * 4 <= g.lib.b2q_streamk_workspace_bytes()
    seen = {}
    owners, contribs, load = {}, {}, []
    for p in range(P):
        _, items = items_of(M, K, N, p)
        units = 0
        sk_items = [it for it in items if it[0] >= dp_tiles]
        for i, (tile, kb0, kb1, role) in enumerate(items):
            assert 0 <= tile < tiles and 0 <= kb0 < kb1 <= nkb
            units += kb1 - kb0
            for kb in range(kb0, kb1):
                assert (tile, kb) not in seen, "k-block processed twice"
                seen[(tile, kb)] = p
            kind = role & 15
            if kb0 == 0 and kb1 == nkb:
                assert role == 0
            elif kb0 == 0:
                assert kind == 1 and tile >= dp_tiles and tile not in owners
                owners[tile] = (p, role >> 4)
                assert (tile, kb0, kb1, role) == sk_items[-1], "the owned piece must close the stream-K segment"
            else:
                assert role == 2 and tile >= dp_tiles
                contribs.setdefault(tile, []).append(p)
                assert i == 0, "a partial must be parked in the pair's first item"
        assert sum(1 for it in items if it[3] == 2) <= 1              # one workspace slot per pair
        # stream-K items come first, then whole data-parallel tiles p, p + P, ...
        assert [it[0] for it in items if it[0] < dp_tiles] == list(range(p, dp_tiles, P))
        assert all(it[0] >= dp_tiles for it in items[:len(sk_items)])
        load.append(units)
    assert len(seen) == tiles * nkb, "k-blocks missed"
    assert set(owners) == set(contribs), "every split tile needs an owner and at least one contributor"
    for tile, (p, expect) in owners.items():
        assert expect == len(contribs[tile]) and min(contribs[tile]) > p
        assert contribs[tile] == sorted(contribs[tile])
    assert max(load) - min(load) <= 1 + (nkb if dp_tiles % P else 0), (min(load), max(load))
    assert g.lib.b2q_debug_gemm_plan(M, K, N, P, (ctypes.c_int * 5)(), (ctypes.c_int * 4)(), 1) == 0  # no such pair


def test_streamk_abi_validation_without_gpu():
    one = ctypes.c_void_p(16)
    assert g.lib.b2q_streamk_workspace_bytes() == 4096 + 74 * 2 * 128 * 256 * 4
    call = lambda M, bits, ws: g.lib.b2q_gemm_streamk(one, one, one, None, None, None, one, M, 4096, 4096, bits, 128, 0,
                                                      None, 0, ws, None)  # noqa: E731
    assert call(128, 4, one) == -2 and b"M > 128" in g.lib.b2q_last_error()
    assert call(2048, 8, one) == -2
    assert call(2048, 4, None) == -2
