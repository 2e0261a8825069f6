"""CPU-side checks (no GPU, no compute calls): the C-ABI library loads and exports every symbol the
header declares, refuses to run without a device, and the panel plan shards a run completely."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    from distance_b200 import api as a
    a.build_library()
    a.load_library()
    return a


def test_header_symbols_are_exported(api):
    hdr = open(os.path.join(ROOT, "include", "distance_gpu.h")).read()
    declared = sorted(set(re.findall(r"DG_API[^;(]*?\b(dg_[a-z0-9_]+)\s*\(", hdr)))
    assert declared == sorted(api.ABI_SYMBOLS)
    lib = api.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.dg_abi_version() == 2


def test_no_cpu_fallback(api):
    """Without a CUDA device the product must fail loudly, never compute on the CPU."""
    if api.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(api.DistanceGpuError) as ei:
        api.Engine("raw", 100)
    assert ei.value.code == -2 and "no CPU fallback" in ei.value.msg


def test_product_does_not_touch_the_oracle():
    """Nothing under distance_b200/ may import, link or load anything under oracle/."""
    for base, _, files in os.walk(os.path.join(ROOT, "distance_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "liboracle" not in txt and "distance_oracle" not in txt, f
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), f


@pytest.mark.parametrize("measure,mode,n_rows,n_cols", [
    ("n_high", 0, 20000, 20000), ("jc69", 0, 100000, 100000), ("tn93", 1, 10000, 10000),
    ("raw", 0, 1000, 1000), ("k80", 0, 2, 2), ("raw", 0, 1, 1), ("n", 1, 3, 70000)])
def test_plan_covers_every_pair_once(api, measure, mode, n_rows, n_cols):
    plan = api.plan_panels(measure, mode, n_rows, n_cols)
    total = n_rows * (n_rows - 1) // 2 if mode == 0 else n_rows * n_cols
    assert sum(p[2] for p in plan) == total
    rows = n_rows - 1 if mode == 0 else n_rows
    if total:
        assert plan[0][0] == 0 and plan[-1][1] == rows
        assert all(plan[k][1] == plan[k + 1][0] for k in range(len(plan) - 1))
    else:
        assert plan == []
    for rb, re_, nr in plan:
        want = sum(n_rows - 1 - i for i in range(rb, re_)) if mode == 0 else (re_ - rb) * n_cols
        assert nr == want


def _rank_main(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from distance_b200 import api, dist
    d = dist.Dist(backend="gloo")
    plan = api.plan_panels("n_high", 0, 20000, 20000, panel_bytes=16 << 20)
    mine = dist.my_panels(plan, d.rank, d.world)
    d.barrier()
    pairs = d.sum(sum(p[2] for p in mine))
    slowest = d.max(10.0 * (rank + 1))
    q.put((rank, len(mine), pairs, slowest))
    d.close()


def test_two_rank_sharding_over_gloo(api):
    """world_size 2 over gloo: ranks take disjoint panels (no data-path collective); the bench
    protocol's SUM of pairs equals the whole triangle and MAX picks the slowest rank."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    total = 20000 * 19999 // 2
    assert out[0][2] == total and out[1][2] == total
    assert out[0][3] == 20.0 and out[1][3] == 20.0
    n_panels = len(api.plan_panels("n_high", 0, 20000, 20000, panel_bytes=16 << 20))
    assert out[0][1] + out[1][1] == n_panels and abs(out[0][1] - out[1][1]) <= 1


def test_pack_nibbles_layout():
    """DG_INPUT_NIBBLE rows: site 2k in the low nibble of byte k, site 2k + 1 in the high one, possibility bits of the
    Paradis code (encoding.rs:7-38); odd widths pad with N (15)."""
    import numpy as np
    from distance_b200 import api
    codes = np.array([[136, 72, 40, 24, 192], [240, 244, 242, 112, 48]], dtype=np.uint8)   # A G C T R / N - ? B Y
    nib = api.pack_nibbles(codes)
    assert nib.shape == (2, 3)
    assert nib.tolist() == [[0x48, 0x12, 0xFC], [0xFF, 0x7F, 0xF3]]


def test_small_panels_fill_whole_rounds_of_the_cta_pairs():
    """The pipelined session cuts the triangle into a dozen panels; each is ONE persistent launch over 74 CTA pairs, so a
    panel of 84 tiles (1.14 rounds) would run at 57 %.  The planner grows small panels up to twice their budget for a
    fuller last round: >= 90 % of the slots of all rounds are busy for config 2's session-sized panels."""
    from distance_b200 import api
    n = 20000
    total_bytes = n * (n - 1) // 2 * 4
    plan = api.plan_panels("n_high", api.DG_MODE_SQUARE, n, n, total_bytes // 8)
    assert sum(p[2] for p in plan) == n * (n - 1) // 2

    def live(r0, r1):   # 512 x 240 tiles right of the diagonal
        blocks = (n + 239) // 240
        return sum(max(0, blocks - (rs + 1) // 240) for rs in range(r0, r1, 512))

    used = sum(live(a, b) for a, b, _ in plan)
    slots = sum(-(-live(a, b) // 74) * 74 for a, b, _ in plan)
    assert used / slots >= 0.90, (used / slots, [b - a for a, b, _ in plan])
