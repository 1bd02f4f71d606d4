"""GPU checks of OPT-IN code paths (not part of `-m gpu`): run with `pytest tests -m gpu_experimental`.

FLIMO_KNN_PAIR=1 — two lanes per query in the first scan (csrc/match_kernel.cu, pair_scan_round).  Validated
bit-identical against the oracle on the headline config c2 (tools/tune_knn.py ... :pair=1 --check) at the end of
round 1; these cases cover what c2 does not (tiny runs, dense cells, caps, empty blocks) and have to be green
before the switch becomes the default."""
import os

import numpy as np
import pytest

from fast_limo_b200 import api, synth

pytestmark = pytest.mark.gpu_experimental
BIG = 1 << 18


def pair_mapper(**kw):
    kw.setdefault("MAX_NUM_MATCHES", BIG)
    kw.setdefault("MAX_NUM_PC2MATCH", BIG)
    os.environ["FLIMO_KNN_PAIR"] = "1"
    try:
        return api.Mapper(api.MappingConfig(**kw), device=0)
    finally:
        del os.environ["FLIMO_KNN_PAIR"]


@pytest.mark.parametrize("name", ["tiny", "c1"])
@pytest.mark.parametrize("cell", [0.0, 0.15, 0.6])
def test_pair_scan_per_point_parity(oracle, flimo_lib, name, cell):
    case = synth.make_case(name)
    m = pair_mapper(knn_cell=cell)
    m.add(case.map_pts, 0.0)
    m.set_scan(case.scan)
    om = oracle.OracleMap()
    om.add(case.map_pts)
    ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4), case.init[:14], case.scan)
    dbg = m.match_debug(case.init)
    assert np.array_equal(dbg["good"], ref["good"])
    g = ref["good"]
    assert np.array_equal(dbg["plane"][g], ref["plane"][g]) and np.array_equal(dbg["dist"][g], ref["dist"][g])
    close = ref["nn_d2"][:, -1] < 2.0
    assert np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close])
    r = m.match(case.init)
    assert r.n_valid == ref["n_valid"] and np.allclose(r.HTH, ref["HTH"], rtol=1e-12, atol=1e-9)


def test_pair_scan_dense_sparse_and_update(oracle, flimo_lib):
    """Very dense cells (runs beyond the private cap), an almost empty map, and the iterated update."""
    rng = np.random.default_rng(1)
    dense = (rng.normal(0, [0.3, 0.3, 0.01], (60000, 3))).astype(np.float32)
    sparse = rng.uniform(-30, 30, (40, 3)).astype(np.float32)
    scan = np.concatenate([rng.normal(0, [0.5, 0.5, 0.02], (3000, 3)), rng.uniform(-30, 30, (500, 3))]).astype(np.float32)
    ident = synth.make_state([0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 1.0])
    for pts in (dense, sparse, np.concatenate([dense, sparse])):
        m = pair_mapper()
        m.add(pts, 0.0)
        m.set_scan(scan)
        om = oracle.OracleMap()
        om.add(pts)
        ref = om.match(oracle.make_cfg(max_pc2match=BIG, max_matches=BIG, num_threads=4), ident[:14], scan)
        dbg = m.match_debug(ident)
        assert np.array_equal(dbg["good"], ref["good"])
        close = ref["nn_d2"][:, -1] < 2.0
        assert np.array_equal(dbg["nn_d2"][close], ref["nn_d2"][close])
    case = synth.make_case("tiny")
    m, m0 = pair_mapper(), api.Mapper(api.MappingConfig(MAX_NUM_MATCHES=BIG, MAX_NUM_PC2MATCH=BIG), device=0)
    for mm in (m, m0):
        mm.add(case.map_pts, 0.0)
        mm.set_scan(case.scan)
    xa, Pa, pa = m.update(case.init, synth.default_P0(), 2, 0.0)
    xb, Pb, pb = m0.update(case.init, synth.default_P0(), 2, 0.0)
    assert pa == pb and np.array_equal(xa, xb) and np.array_equal(Pa, Pb)
