"""CPU: the oracle's octree (oracle/ioctree.hpp) against the REFERENCE's own octree.

oracle/_ref/libref_octree.so is the reference's fast_limo/Objects/Octree.hpp compiled where it lies, unmodified
(`make -C oracle ref`; its only dependency, Eigen::Vector3f, is supplied by the stand-in oracle/ref_shim).
tests/golden/ref_octree_*.npz hold outputs of THAT library on seeded inputs (tests/golden/make_golden.py), so
these checks run wherever the fixtures are — also where /root/reference is not mounted: map size after every
Octree::update (leaf splits, batch-atomic down-sampling, root growth, NaN points), final contents, exact 5-NN.
"""
import importlib.util
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(G, "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_oracle_octree_matches_reference_golden(oracle, tag):
    mg = _mg()
    g = np.load(os.path.join(G, f"ref_octree_{tag}.npz"))
    batches, queries = mg.ref_octree_inputs()
    assert sum(float(np.nansum(b.astype(np.float64))) for b in batches) == float(g["in_checksum"])
    om = oracle.OracleMap(min_extent=float(g["min_extent"]), downsample=bool(g["downsample"]))
    sizes = []
    for b in batches:
        om.add(b)
        sizes.append(om.size())
    assert sizes == g["sizes"].tolist()
    assert np.array_equal(mg.contents_checksum(om.points()), g["contents_checksum"])
    d2, nb, cnt = om.knn(queries, 5)
    assert np.array_equal(cnt, g["knn_cnt"]) and np.array_equal(d2, g["knn_d2"])          # bit-exact, same order
    assert float(nb.astype(np.float64).sum()) == float(g["knn_xyz_checksum"])


def test_oracle_octree_matches_live_reference(oracle):
    """Same comparison against the compiled reference itself, on more inputs (only where oracle/_ref exists)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_octree.so not built (needs /root/reference)")
    rng = np.random.default_rng(99)
    for me, ds in ((0.2, True), (0.5, True), (0.2, False)):
        om = oracle.OracleMap(min_extent=me, downsample=ds)
        ro = oracle.RefOctree(bucket=2, min_extent=me, downsample=ds)
        centres = rng.uniform(-20, 20, (4, 3)) * [1, 1, 0.1]
        for i in range(12):
            c = centres[i % 4]
            b = (c + rng.normal(0, [2.0, 2.0, 0.05], (3000, 3))).astype(np.float32)
            if i == 5:
                b += np.float32(150.0)
            om.add(b)
            ro.add(b)
            assert om.size() == ro.size()
        a, r = om.points(), ro.points()
        assert np.array_equal(a[np.lexsort(a.T)], r[np.lexsort(r.T)])
        q = (centres[rng.integers(4, size=4000)] + rng.normal(0, 2.5, (4000, 3))).astype(np.float32)
        d2, nb, cnt = om.knn(q, 5)
        rx, rd2, rcnt = ro.knn(q, 5)
        assert np.array_equal(d2, rd2) and np.array_equal(nb, rx) and np.array_equal(cnt, rcnt)
    # Mapper::set_config's setBucketSize is a self-assignment (Octree.hpp:178-180): the YAML value changes nothing
    r2, r32 = oracle.RefOctree(bucket=2), oracle.RefOctree(bucket=32)
    pts = rng.normal(0, 3, (20000, 3)).astype(np.float32)
    r2.add(pts)
    r32.add(pts)
    assert np.array_equal(r2.knn(pts[:500], 5)[1], r32.knn(pts[:500], 5)[1])


def test_property_random_update_sequences_match_reference(oracle):
    """Property test (hypothesis): random sequences of batches — clustered, duplicated, far away, with NaNs,
    tiny first scans — give the same tree population and the same exact kNN in the restatement and in the
    compiled reference octree."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libref_octree.so not built (needs /root/reference)")
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(seed=st.integers(0, 2**31 - 1), n_batches=st.integers(1, 7), min_extent=st.sampled_from([0.1, 0.2, 0.35, 1.0]),
           downsample=st.booleans(), spread=st.sampled_from([0.3, 2.0, 15.0]))
    def run(seed, n_batches, min_extent, downsample, spread):
        rng = np.random.default_rng(seed)
        om = oracle.OracleMap(min_extent=min_extent, downsample=downsample)
        ro = oracle.RefOctree(bucket=2, min_extent=min_extent, downsample=downsample)
        centres = rng.uniform(-10, 10, (3, 3))
        allp = []
        for b in range(n_batches):
            n = int(rng.integers(1, 1500))
            p = (centres[rng.integers(3, size=n)] + rng.normal(0, spread, (n, 3))).astype(np.float32)
            if rng.random() < 0.2:
                p[rng.integers(n)] = np.nan
            if rng.random() < 0.15:
                p = (p + np.float32(rng.uniform(50, 400))).astype(np.float32)
            if rng.random() < 0.2 and allp:
                p = np.concatenate([p, allp[-1][: len(allp[-1]) // 2]])            # exact duplicates of earlier points
            om.add(p)
            ro.add(p)
            allp.append(p)
            assert om.size() == ro.size()
        a, r = om.points(), ro.points()
        assert np.array_equal(a[np.lexsort(a.T)], r[np.lexsort(r.T)])
        q = (centres[rng.integers(3, size=300)] + rng.normal(0, spread * 1.5, (300, 3))).astype(np.float32)
        d2, nb, cnt = om.knn(q, 5)
        rx, rd2, rcnt = ro.knn(q, 5)
        assert np.array_equal(cnt, rcnt) and np.array_equal(d2, rd2) and np.array_equal(nb, rx)

    run()
