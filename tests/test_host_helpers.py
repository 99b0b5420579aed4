"""CPU tests of host-side helpers that mirror reference utilities: SectionTimer
(tools/timer.py), element adjacency (mesh/utils.py:139-263, fea/composer.py:351-371),
task statistics and the small argparse converters."""
import numpy as np
import pytest


def test_section_timer_interface():
    import sktopt
    from sktopt.tools.timer import SectionStats, SectionTimer
    clock = iter(np.arange(0.0, 100.0, 1.0))
    t = SectionTimer(clock=lambda: float(next(clock)), hierarchical=True)
    with t.section("outer"):            # start 0
        with t.section("inner"):        # start 1, end 2
            pass
        with t.section("inner"):        # start 3, end 4
            pass
    # outer ends at 5
    t.add("manual", 0.25)

    @t.wrap("wrapped")
    def f(x):
        return x + 1
    assert f(1) == 2 and f.__name__ == "f"
    st = {s.name: s for s in t.stats()}
    assert isinstance(st["outer"], SectionStats)
    assert st["outer"].total == 5.0 and st["outer"].count == 1
    assert st["outer>inner"].total == 2.0 and st["outer>inner"].count == 2
    assert st["outer>inner"].avg == 1.0 and st["outer>inner"].max == 1.0
    assert st["manual"].total == 0.25 and st["wrapped"].count == 1
    assert [s.name for s in t.summary()][0] == "outer"
    assert [s.name for s in t.summary(sort_by="name", descending=False)][0] == "manual"
    own = {s.name: s.total for s in t.summary_self_time()}
    assert own["outer"] == 3.0 and own["outer>inner"] == 2.0
    with pytest.raises(ValueError):
        t.summary(sort_by="bogus")
    text = t.report()
    assert text.splitlines()[0].startswith("outer: total=5.000000s")
    with pytest.raises(RuntimeError, match="matplotlib"):
        t.plot_bar()
    t.reset("manual")
    assert "manual" not in {s.name for s in t.stats()}
    t.reset()
    assert t.stats() == [] and t.report() == "No timing data collected."
    with pytest.raises(ValueError, match="No timing data"):
        t.plot()
    flat = sktopt.tools.SectionTimer()
    with flat.section("a"):
        with flat.section("b"):
            pass
    assert {s.name for s in flat.stats()} == {"a", "b"}


def test_element_adjacency_matches_bruteforce():
    import sktopt
    from sktopt.mesh import utils
    from sktopt.fea import composer
    mesh = sktopt.mesh.toy_problem.create_box_hex(2.0, 1.5, 1.0, 0.5)
    ne = mesh.nelements
    sets = [set(mesh.t[:, e].tolist()) for e in range(ne)]
    brute = np.array([[1 if sets[i] & sets[j] else 0 for j in range(ne)] for i in range(ne)])
    A = utils.build_element_adjacency_matrix(mesh)
    assert A.dtype == np.uint8 and np.array_equal(A.toarray(), brute)
    B = utils.build_element_adjacency_matrix_fast(mesh)
    assert np.array_equal(B.toarray(), brute - np.eye(ne, dtype=int))
    nb = utils.get_adjacent_elements_fast(B, [0, 1])
    expect = sorted(set(np.nonzero(brute[0] | brute[1])[0].tolist()) - {0, 1})
    assert nb.dtype == np.int32 and nb.tolist() == expect
    assert nb.tolist() == utils.get_adjacent_elements(mesh, [0, 1])
    # face neighbours of tetrahedra
    tet = sktopt.mesh.toy_problem.create_box_tet(1.0, 1.0, 1.0, 0.5)
    adj = composer.adjacency_matrix(tet)
    faces = [{frozenset(c) for c in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))}]
    for e in range(tet.nelements):
        fe = {frozenset(tet.t[list(c), e].tolist()) for c in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))}
        expect = sorted(j for j in range(tet.nelements) if j != e and fe & {
            frozenset(tet.t[list(c), j].tolist()) for c in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3))})
        assert sorted(adj[e]) == expect


def test_task_statistics_and_converters(capsys):
    import sktopt
    from sktopt.core import misc
    tsk = sktopt.mesh.toy_problem.toy_test()
    st = tsk.nodes_and_elements_stats()
    assert abs(st["nodes"]["min"] - 1.0) < 1e-12 and abs(st["elements"]["max"] - 1.0) < 1e-12
    assert "=== Distance between nodes ===" in capsys.readouterr().out
    assert np.array_equal(tsk.force_elements_all, np.unique(tsk.neumann_elements))
    t2 = sktopt.mesh.toy_problem.toy2()
    assert t2.n_tasks == 2
    assert np.array_equal(t2.force_elements_all,
                          np.unique(np.concatenate([np.ravel(a) for a in t2.neumann_elements])))
    assert misc.str2bool("Yes") is True and misc.str2bool("0") is False and misc.str2bool(True) is True
    with pytest.raises(Exception):
        misc.str2bool("maybe")
    assert misc.float_or_none("None") is None and misc.float_or_none("1.5") == 1.5


def test_oc_bisection_saturation_predicate_is_exact():
    """``rates_all_on_lower_clip`` (core/optimizers/oc.py) lets the device bisection
    skip midpoints whose candidate cannot differ from the one just evaluated.  Against
    the reference's candidate formula (oc.py:50-68, restated in NumPy): whenever the
    predicate is True for two midpoints, the two candidates are bit-identical; and over
    the reference's bracket [1e-7, 1e7] it fires for the first 10-20 midpoints."""
    pytest.importorskip("torch")
    from sktopt.core.optimizers.oc import rates_all_on_lower_clip
    rng = np.random.default_rng(0)
    n = 4000
    eps, move, rmin, rmax, smin, smax = 1e-12, 0.2, 1e-3, 1.0, 0.7, 1.3

    def candidate(dC, rho, lmid, eta):
        with np.errstate(invalid="ignore"):
            sr = np.clip(np.power(-dC / (lmid + eps), eta), smin, smax)
        lo = np.maximum(rho - move, rmin)
        hi = np.minimum(rho + move, rmax)
        return np.clip(rho * sr, lo, hi)

    for eta in (0.5, 0.3, 1.0):
        dC = -rng.lognormal(0.0, 2.0, n)            # sensitivities are negative
        dC[::97] = 0.0
        rho = rng.uniform(rmin, rmax, n)
        neg_max = float(np.max(-dC))
        l1, l2, fired, ref = 1e-7, 1e7, 0, None
        for _ in range(60):
            lmid = 0.5 * (l1 + l2)
            if rates_all_on_lower_clip(neg_max, lmid, eps, eta, smin):
                fired += 1
                c = candidate(dC, rho, lmid, eta)
                if ref is None:
                    ref = c
                assert np.array_equal(c, ref)       # bit-identical, not just close
                assert np.array_equal(c, np.clip(rho * smin, np.maximum(rho - move, rmin),
                                                 np.minimum(rho + move, rmax)))
            l2 = lmid                               # walk down, as while the volume is short
        assert fired >= 8, (eta, fired)      # 5e6 -> max(-dC) / smin^(1/eta) in halvings
        # just inside / outside the boundary the predicate stays on the safe side
        lam_star = neg_max / smin ** (1.0 / eta) - eps
        assert not rates_all_on_lower_clip(neg_max, lam_star, eps, eta, smin)
        assert not rates_all_on_lower_clip(neg_max, lam_star * (1 - 1e-6), eps, eta, smin)
        assert rates_all_on_lower_clip(neg_max, lam_star * (1 + 1e-6), eps, eta, smin)
    # degenerate inputs never skip
    assert not rates_all_on_lower_clip(-1.0, 1.0, eps, 0.5, smin)
    assert not rates_all_on_lower_clip(float("nan"), 1.0, eps, 0.5, smin)
    assert not rates_all_on_lower_clip(1.0, 1.0, eps, 0.0, smin)


def test_lattice_facet_table_equals_the_sort_based_one():
    """``MeshHex._build_facets_lattice`` (O(n) closed form for lattice-numbered grids)
    must reproduce the generic lexicographic construction exactly -- facets, t2f, f2t,
    f2lf, dtypes included -- for any geometry and any valid local vertex order, and
    must decline everything else (permuted numbering, one-cell-thick grids)."""
    from sktopt._fem.mesh import Mesh, MeshHex
    rng = np.random.default_rng(0)
    keys = ("_facets", "_t2f", "_f2t", "_f2lf")

    def generic(mesh):
        g = MeshHex(mesh.p, mesh.t)
        Mesh._build_facets(g)
        return g

    mirror = np.array([1, 0, 4, 5, 2, 3, 7, 6])       # reflect the reference cube along Z
    for dims in ((4, 5, 3), (7, 4, 9), (3, 3, 3), (12, 9, 8), (5, 4, 2)):
        m = MeshHex.init_tensor(*[np.linspace(0, 1, d) for d in dims])
        p = m.p + rng.uniform(-0.01, 0.01, m.p.shape)
        t2 = m.t.copy()
        flip = rng.uniform(size=t2.shape[1]) < 0.5
        t2[:, flip] = t2[mirror][:, flip]
        for t in (m.t, t2):
            a = MeshHex(p, t)
            assert a._build_facets_lattice(), dims
            g = generic(a)
            for k in keys:
                assert np.array_equal(getattr(a, k), getattr(g, k)), (dims, k)
                assert getattr(a, k).dtype == getattr(g, k).dtype, (dims, k)
            # the public properties go through the fast path
            b = MeshHex(p, t)
            assert np.array_equal(b.facets, g._facets) and np.array_equal(b.f2t, g._f2t)
            assert np.array_equal(b.boundary_facets(), np.nonzero(g._f2t[1] == -1)[0])
    m = MeshHex.init_tensor(*[np.linspace(0, 1, d) for d in (5, 4, 6)])
    perm = rng.permutation(m.p.shape[1])
    scrambled = MeshHex(m.p[:, np.argsort(perm)], perm[m.t].astype(np.int32))
    assert not scrambled._build_facets_lattice()
    g = generic(scrambled)
    assert np.array_equal(scrambled.facets, g._facets)       # falls back to the generic path
    for dims in ((2, 5, 4), (5, 2, 4)):
        thin = MeshHex.init_tensor(*[np.linspace(0, 1, d) for d in dims])
        assert not thin._build_facets_lattice()
        assert np.array_equal(thin.facets, generic(thin)._facets)


def test_native_host_helpers_reproduce_numpy_bit_for_bit():
    """csrc/host_setup.cu (multi-threaded C++ on HOST pointers, no device work): the
    hexahedral element volumes and the lattice facet table must equal the NumPy
    constructions bit for bit, on jittered geometry, mirrored local vertex orders and a
    mesh large enough to run on several threads; meshes without the lattice structure
    are declined."""
    from sktopt._b200 import lib as _lib
    try:
        _lib.load()
    except RuntimeError:
        pytest.skip("library not built")
    from sktopt._fem.mesh import Mesh, MeshHex
    from sktopt.fea import composer
    rng = np.random.default_rng(1)
    mirror = np.array([1, 0, 4, 5, 2, 3, 7, 6])
    for dims in ((4, 5, 3), (7, 4, 9), (3, 3, 3), (5, 4, 2), (70, 60, 45)):
        m = MeshHex.init_tensor(*[np.linspace(0, 1, d) for d in dims])
        p = m.p + rng.uniform(-0.002, 0.002, m.p.shape)
        t2 = m.t.copy()
        flip = rng.uniform(size=t2.shape[1]) < 0.5
        t2[:, flip] = t2[mirror][:, flip]
        for t in (m.t, t2):
            a = MeshHex(p, t)
            assert a._build_facets_native(), dims
            g = MeshHex(p, t)
            Mesh._build_facets(g)
            for k in ("_facets", "_t2f", "_f2t", "_f2lf"):
                assert np.array_equal(getattr(a, k), getattr(g, k)), (dims, k)
                assert getattr(a, k).dtype == getattr(g, k).dtype, (dims, k)
            v_nat = composer._hex_volumes_native(t, p)
            used = sorted({k for quad in composer._HEX_TETS for k in quad})
            P = {k: p[:, t[k]] for k in used}
            v_np = np.zeros(t.shape[1])
            for quad in composer._HEX_TETS:
                v_np += composer._abs_tet_volume(P, quad)
            assert np.array_equal(v_nat, v_np), dims
            assert np.array_equal(composer.get_elements_volume(a), v_np)
    perm = rng.permutation(m.p.shape[1])
    scrambled = MeshHex(m.p[:, np.argsort(perm)], perm[m.t].astype(np.int32))
    assert not scrambled._build_facets_native()
    thin = MeshHex.init_tensor(*[np.linspace(0, 1, d) for d in (2, 5, 4)])
    assert not thin._build_facets_native()


def test_mean_of_vertices_equals_numpy_mean_bitwise():
    from sktopt._fem.mesh import _mean_of_vertices
    rng = np.random.default_rng(2)
    p = rng.standard_normal((3, 5000)) * np.array([[1.0], [1e3], [1e-3]])
    for k in (3, 4, 8):
        conn = rng.integers(0, p.shape[1], (k, 20000)).astype(np.int32)
        assert np.array_equal(_mean_of_vertices(p, conn), p[:, conn].mean(axis=1))
