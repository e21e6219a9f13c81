"""CPU tests of the product's host side (no GPU, no compute): the C-ABI library loads and exports every
declared symbol, and the compiled structures equal the oracle's integer for integer -- junction tree
(elimination order, clusters, separators, factor assignment), per-message bucket-elimination traces
("clique index maps and elimination order bit-exact"), VE orders and traces."""
import ctypes
import os
import re

import numpy as np
import pytest

import hbayes_b200 as hb
from hbayes_b200 import _lib, synth
from oracle import OracleJT, OracleNet
from helpers import G, GOLDEN_NETS, oracle_traces, product_traces, random_evidence_list, random_net

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "hbayes_cuda.h")).read()
    declared = set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_no_cpu_fallback():
    """Without a device a compile call fails with HB_E_CUDA; a structure-only handle refuses to compute."""
    net = hb.importBayesianGraph(os.path.join(G, "cancer.net"))
    jt = hb.JunctionTree(net, device=None)
    with pytest.raises(hb.HbError) as e:
        jt.posteriors_batch(np.full((2, net.n), -1, np.int8))
    assert e.value.code == _lib.HB_E_CUDA
    if _lib.lib().hb_device_count() == 0:
        with pytest.raises(hb.HbError) as e:
            hb.JunctionTree(net, device=0)
        assert e.value.code == _lib.HB_E_CUDA


def test_hugin_loader_matches_oracle():
    for f in GOLDEN_NETS:
        a = hb.importBayesianGraph(os.path.join(G, f))
        b = OracleNet.load_hugin(os.path.join(G, f))
        assert (a.names, a.dims, a.scopes, a.values) == (b.names, b.dims, b.scopes, b.values)
    assert hb.importBayesianGraph(os.path.join(G, "missing.net")) is None


def test_gui_export_dialect_loads_like_the_plain_file():
    """a Hugin / SamIam GUI export -- `net { HR_... }` property header, node attributes, tab-indented nested data blocks, the
    dialect of the reference's bundled cancer.net (/root/reference/cancer.net:1-128) -- gives exactly the network of the plain
    file, in the product's loader and in the oracle's"""
    plain = hb.importBayesianGraph(os.path.join(G, "sprinkler.net"))
    for loader in (hb.importBayesianGraph, OracleNet.load_hugin):
        gui = loader(os.path.join(G, "sprinkler_gui.net"))
        assert (gui.names, gui.dims, gui.scopes, gui.values) == (plain.names, plain.dims, plain.scopes, plain.values)
    text = open(os.path.join(G, "sprinkler_gui.net")).read()
    assert "HR_Compile_TriangMethod" in text and "\tposition = (" in text and "data = (((" in text


def test_bad_arguments_are_status_codes():
    with pytest.raises(hb.HbError) as e:
        hb.BayesianNetwork.from_tables([2, 2], [[0], [1, 0]], [[0.5, 0.5], [0.1, 0.9, 0.3]])
    assert e.value.code == _lib.HB_E_ARG
    net = hb.importBayesianGraph(os.path.join(G, "asia.net"))
    jt = hb.JunctionTree(net, device=None)
    with pytest.raises(hb.HbError):
        jt.program([0, 0])  # observed twice
    with pytest.raises(hb.HbError):
        jt.program([99])


def check_tree_and_traces(pnet, onet, rng, trials, max_obs):
    pjt = hb.JunctionTree(pnet, device=None)
    ojt = OracleJT(onet)
    assert pjt.describe() == ojt.describe()
    for _ in range(trials):
        ev = random_evidence_list(rng, onet.dims, max_obs)
        want = oracle_traces(ojt, ev)
        prog = pjt.program([v for v, _ in ev])
        got = product_traces(prog)
        assert got.keys() == want.keys()
        for k in want:
            assert got[k] == want[k], (ev, k)


@pytest.mark.parametrize("fname", GOLDEN_NETS)
def test_golden_nets_structure_and_traces(fname):
    pnet = hb.importBayesianGraph(os.path.join(G, fname))
    onet = OracleNet.load_hugin(os.path.join(G, fname))
    check_tree_and_traces(pnet, onet, np.random.RandomState(1), trials=8, max_obs=4)


@pytest.mark.parametrize("seed", range(30))
def test_random_nets_structure_and_traces(seed):
    net = random_net(seed, n=4 + seed % 12, states=(2, 4), max_parents=3, shuffle=seed % 2 == 1)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    check_tree_and_traces(pnet, onet, np.random.RandomState(seed), trials=4, max_obs=5)


def test_added_edges_comparator_and_explicit_order():
    net = random_net(5, n=12, states=(2, 3), max_parents=3)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    assert hb.JunctionTree(pnet, cmp=hb.numberOfAddedEdges, device=None).describe() == OracleJT(onet, cmp="added").describe()
    order = np.random.RandomState(3).permutation(net.n).tolist()
    assert hb.JunctionTree(pnet, order=order, device=None).describe() == OracleJT(onet, order=order).describe()


def test_c2_structure_matches_oracle():
    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    onet = OracleNet.from_arrays(net.dims, net.scopes, net.values)
    pjt = hb.JunctionTree(pnet, device=None)
    ojt = OracleJT(onet)
    assert pjt.describe() == ojt.describe()
    ev = [(v, 0) for v in observed]
    assert product_traces(pjt.program(observed)) == oracle_traces(ojt, ev)


@pytest.mark.parametrize("shape", [(3, 3), (4, 5), (6, 6)])
def test_ve_orders_and_traces(shape):
    g = synth.grid_net(shape[0], shape[1], 11)
    pnet = hb.BayesianNetwork.from_tables(g.dims, g.scopes, g.values)
    onet = OracleNet.from_arrays(g.dims, g.scopes, g.values)
    assert hb.minFillOrder(pnet) == onet.ve_order("min_fill")
    assert hb.minDegreeOrder(pnet) == onet.ve_order("min_degree")
    r = [g.n - 1]
    p = [v for v in hb.minFillOrder(pnet) if v not in r]
    ve = hb.VariableElimination(pnet, p, r, device=None)
    rng = np.random.RandomState(0)
    for _ in range(4):
        ev = random_evidence_list(rng, g.dims, 4)
        ev = [(v, s) for v, s in ev if v not in r]
        _, _, want = onet.posterior_marginal(p, r, ev, trace=True)
        got = ve.program([v for v, _ in ev])["groups"][0]["steps"]
        assert got == want


def _tree_properties(d, n_vars):
    """Bayes/FactorElimination.hs:341-378: every variable's clusters form a connected subtree (running
    intersection), separators are the intersections of their end clusters, every CPT family sits in its cluster."""
    clusters = [set(c) for c in d["clusters"]]
    nc = len(clusters)
    parent = [-1] * nc
    for s, (p, c) in enumerate(zip(d["sep_parent"], d["sep_child"])):
        parent[c] = p
        assert set(d["sep_vars"][s]) == clusters[p] & clusters[c]
    assert sum(1 for p in parent if p < 0) == 1 and parent[d["root"]] == -1
    for v in range(n_vars):
        holders = [c for c in range(nc) if v in clusters[c]]
        assert holders, "variable in no cluster"
        # connected subtree <=> exactly one holder whose parent is not a holder
        tops = [c for c in holders if parent[c] < 0 or v not in clusters[parent[c]]]
        assert len(tops) == 1, (v, holders)
    for c in range(nc):  # maximal: no cluster contains another
        assert not any(c != o and clusters[c] <= clusters[o] for o in range(nc))
    return clusters


@pytest.mark.parametrize("seed", range(12))
def test_running_intersection_and_family_coverage(seed):
    net = random_net(100 + seed, n=6 + 3 * seed, states=(2, 4), max_parents=3, shuffle=seed % 3 == 0)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    for cmp in (hb.nodeComparisonForTriangulation, hb.numberOfAddedEdges):
        d = hb.JunctionTree(pnet, cmp=cmp, device=None).describe()
        clusters = _tree_properties(d, net.n)
        placed = sorted(v for f in d["factors"] for v in f)
        assert placed == list(range(net.n))  # every CPT assigned exactly once
        for c, fs in enumerate(d["factors"]):
            for v in fs:
                assert set(net.scopes[v]) <= clusters[c]


def test_c3_tree_properties():
    net, _ = synth.config_c3()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    _tree_properties(hb.JunctionTree(pnet, device=None).describe(), net.n)


def test_plain_c_program_uses_the_abi(tmp_path):
    import subprocess

    exe = str(tmp_path / "c_abi_smoke")
    src = os.path.join(ROOT, "tests", "c_abi_smoke.c")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", src, "-o", exe, "-L" + libdir, "-lhbayes_cuda", "-Wl,-rpath," + libdir],
                   check=True)
    r = subprocess.run([exe, os.path.join(G, "cancer.net")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert '"clusters":[[2,3,4],[0,3,4],[1,3]]' in r.stdout


def test_save_and_load_round_trip(tmp_path):
    net, observed = synth.config_c2()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.JunctionTree(pnet, device=None)
    path = tmp_path / "c2.hbjt"
    jt.save(path)
    back = hb.JunctionTree.load(path, pnet, device=None)
    assert back.describe() == jt.describe()
    assert back.program(observed) == jt.program(observed)
    (tmp_path / "junk").write_bytes(b"not a tree")
    with pytest.raises(hb.HbError) as e:
        hb.JunctionTree.load(tmp_path / "junk", pnet, device=None)
    assert e.value.code == _lib.HB_E_IO


def test_degree_order_matches_twin():
    from oracle import twin

    for shape in [(3, 3), (4, 6)]:
        g = synth.grid_net(shape[0], shape[1], 3)
        pnet = hb.BayesianNetwork.from_tables(g.dims, g.scopes, g.values)
        tnet = twin.Network.from_tables(g.dims, g.parents, [v.tolist() for v in g.values])
        for order in (hb.minFillOrder(pnet), hb.minDegreeOrder(pnet), list(range(g.n))):
            assert hb.degreeOrder(pnet, order) == twin.degree_order(tnet, order)
    assert hb.degreeOrder(pnet, list(range(g.n))) >= 2


def test_mpe_instantiation_order_is_structural_and_matches_the_twin():
    """the variable order of the reference's instantiation list (ReferencePatterns.hs:79-84 compares it literally) is
    known at compile time: hb_ve_mpe_order on a structure-only handle == the twin's evaluated list"""
    from oracle import twin

    path = os.path.join(G, "asia.net")
    pnet = hb.importBayesianGraph(path)
    tnet = twin.load_hugin(path)
    i = pnet.ids
    x, b, d, a, s, l, t, e = [i[c] for c in "XBDASLTE"]
    for p, r in (([x, d], [b, a, s, l, t, e]), ([x, d, b, l, t, e], [a, s]), ([d, x], [e, t, l, s, b, a])):
        h = hb.MostProbableExplanation(pnet, p, r, device=None)
        _, m = twin.mpe(tnet, p, r, [(x, 0), (d, 1)])
        assert h.order([x, d]) == [v for v, _ in m[0]]
    h = hb.MostProbableExplanation(pnet, [x, d], [b, a, s, l, t, e], device=None)
    assert [pnet.names[v] for v in h.order([x, d])] == ["E", "T", "L", "S", "B", "A"]
    with pytest.raises(hb.HbError):  # evidence on a maximised variable: not supported (see compile_mpe)
        h.order([x, b])
    # a variable in neither list: its last table is dropped by addBucket (Buckets.hs:92-98), in the twin and here alike
    # (the dropped table carries every instantiation with it: the reference answers [] here)
    value, m = twin.mpe(tnet, [x], [b, a, s, l, t, e], [(x, 0)])
    assert m == [] and value == 1.0
    assert hb.MostProbableExplanation(pnet, [x], [b, a, s, l, t, e], device=None).order([x]) == []


def test_specialised_step_kernels_source_compiles_for_sm_100a(tmp_path):
    """csrc/jit.cpp: the per-step CUDA source the library hands to NVRTC on the first large batch -- one kernel per
    eligible step, deterministic, and it compiles for sm_100a (nvcc here; numerics are checked on the GPU)"""
    import shutil
    import subprocess

    net = random_net(11, n=10, states=(2, 3), max_parents=3)
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    jt = hb.JunctionTree(pnet, device=None)
    src = jt.jit_source([1, 4])
    assert src == hb.JunctionTree(pnet, device=None).jit_source([1, 4])
    prog = jt.program([1, 4])
    eligible = [i for i, s in enumerate(prog["physical"])
                if s["kind"] == 0 and not s["shared"] and not s["scalars"] and sum(len(l["in"]) for l in s["levels"]) >= 1]
    assert [int(x.split("(")[0]) for x in src.split("hb_step_")[1:]] == eligible
    assert "__dmul_rn" in src and "__dadd_rn" in src and "fma" not in src
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    cu = tmp_path / "steps.cu"
    cu.write_text(src)
    r = subprocess.run([nvcc, "--fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", str(cu), "-o", str(tmp_path / "steps.cubin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]


def test_uncalibrated_handles_fail_loudly():
    """posterior [v] on a handle that was never calibrated (structure only) is an error, not a read out of bounds"""
    net = hb.importBayesianGraph(os.path.join(G, "cancer.net"))
    jt = hb.JunctionTree(net, device=None)
    with pytest.raises(hb.HbError) as e:
        jt.posterior([0])
    assert e.value.code == _lib.HB_E_CUDA
    with pytest.raises(hb.HbError) as e:
        jt.posterior([0, 3])
    assert e.value.code == _lib.HB_E_CUDA


def test_corrupt_tree_files_are_rejected(tmp_path):
    """every index of a loaded tree is range-checked -- a damaged file gives HB_E_IO, never a crash"""
    import struct

    net = hb.importBayesianGraph(os.path.join(G, "asia.net"))
    jt = hb.JunctionTree(net, device=None)
    good = tmp_path / "asia.hbjt"
    jt.save(good)
    data = good.read_bytes()
    assert hb.JunctionTree.load(good, net, device=None).describe() == jt.describe()
    # the tree section is the tail of the file: damage every 8-byte word of the last 1.5 KB in turn
    start = max(8, len(data) - 1536) // 8 * 8
    rejected = 0
    for at in range(start, len(data) - 7, 8):
        for value in (1 << 30, -7, 3):
            if struct.unpack_from("<q", data, at)[0] == value:
                continue
            bad = bytearray(data)
            struct.pack_into("<q", bad, at, value)
            p = tmp_path / "bad.hbjt"
            p.write_bytes(bytes(bad))
            try:
                t = hb.JunctionTree.load(p, net, device=None)
                t.program([0])  # whatever still loads must be a usable tree
            except hb.HbError as e:
                assert e.code in (_lib.HB_E_IO, _lib.HB_E_ARG), e
                rejected += 1
    assert rejected > 100
    for cut in (len(data) - 8, len(data) // 2, 20):
        p = tmp_path / "short.hbjt"
        p.write_bytes(data[:cut])
        with pytest.raises(hb.HbError) as e:
            hb.JunctionTree.load(p, net, device=None)
        assert e.value.code == _lib.HB_E_IO


def test_on_chip_kernel_source_compiles_for_sm_100a(tmp_path):
    """csrc/jit.cpp hb_tree: one kernel for changeEvidence + collect + distribute + all posteriors, per-row
    tables in shared memory.  Deterministic source, every per-row step present, and it compiles for sm_100a with the
    shared-memory loads the design relies on (nvcc here; numerics are checked on the GPU, tests/test_gpu_onchip.py)."""
    import shutil
    import subprocess

    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    for name, pnet, observed in (("cancer", hb.importBayesianGraph(os.path.join(G, "cancer.net")), [0]),
                                 ("rand", None, [1, 4])):
        if pnet is None:
            net = random_net(11, n=10, states=(2, 3), max_parents=3)
            pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
        jt = hb.JunctionTree(pnet, device=None)
        src = jt.tree_source(observed)
        assert src == hb.JunctionTree(pnet, device=None).tree_source(observed)
        prog = jt.program(observed)
        row_steps = [i for i, s in enumerate(prog["physical"]) if s["kind"] == 0 and not s["shared"]]
        assert [int(x.split(":")[0]) for x in src.split("// step ")[1:]] and \
            sorted(int(x.split(":")[0]) for x in src.split("// step ")[1:]) == row_steps
        assert "hb_tree" in src and "__syncthreads" in src and "__dmul_rn" in src and "__dadd_rn" in src
        if not os.path.exists(nvcc):
            continue
        cu = tmp_path / (name + ".cu")
        cu.write_text(src)
        cubin = tmp_path / (name + ".cubin")
        r = subprocess.run([nvcc, "--fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", str(cu), "-o", str(cubin)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
        # (the only DFMAs are the Newton steps of the correctly rounded 1.0 / norm)
        assert "LDS" in sass and "DMUL" in sass and "DADD" in sass and sass.count("DFMA") < sass.count("DMUL")
    # a program whose per-row tables cannot fit shared memory has no on-chip kernel: the call says so
    net, observed = synth.config_c3()
    cnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    with pytest.raises(hb.HbError) as e:
        hb.JunctionTree(cnet, device=None).tree_source(observed)
    assert e.value.code == _lib.HB_E_UNSUPPORTED


def test_split_program_exchange_steps_are_structural():
    """the program compiled for one slice of a split clique marks every table that holds a partial sum over the
    split variable (messages leaving the split clique, posteriors inside it) for an all-reduce and every query that
    mentions the split variable for a gather -- decided on the host, no GPU or communicator needed"""
    from test_gpu_split_nccl import test_split_program_has_its_exchange_steps

    test_split_program_has_its_exchange_steps()


def _kernels(src):
    """{step: source} of the `hb_step_<i>` kernels in a generated translation unit"""
    parts = re.split(r'(?=extern "C" __global__)', src)[1:]
    return {int(re.search(r"hb_step_(\d+)\(", p).group(1)): p for p in parts}


def test_rereading_steps_get_register_blocks_over_several_output_variables(tmp_path):
    """csrc/jit.cpp jit_step_shape: a step that enumerates a big joint from small per-row tables (the messages around C3's
    largest cluster) evaluates the states of SEVERAL output variables in lock step -- an outer-product block, every loaded
    value used several times -- with one row per thread and the registers of a whole SM partition; the variables its largest
    input does not mention vary fastest (an inner loop where that is worth it).  The choice is deterministic, HB_JIT_BLOCK_U=0
    turns it off, and the blocked kernels compile for sm_100a without spilling much (numerics: tests/test_gpu_configs.py)."""
    import shutil
    import subprocess

    net, observed = synth.config_c3()
    pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)
    src = hb.JunctionTree(pnet, device=None).jit_source(sorted(observed))
    ks = _kernels(src)
    blocked = {i: k for i, k in ks.items() if "__launch_bounds__(256, 1)" in k}
    assert 4 <= len(blocked) <= 40, sorted(blocked)
    prog = hb.JunctionTree(pnet, device=None).program(sorted(observed))
    heaviest = max((i for i, s in enumerate(prog["physical"]) if s["kind"] == 0 and not s["shared"]),
                   key=lambda i: prog["physical"][i]["n_out"] * int(np.prod([l["d"] for l in prog["physical"][i]["levels"]])))
    assert heaviest in blocked
    for i, k in blocked.items():
        assert "double2" not in k.split("{", 1)[1], i  # one row per thread
        st = prog["physical"][i]
        stores = k.count("__stcs(")
        assert stores > st["U"] and st["n_out"] % stores == 0, (i, stores)  # more entries in lock step than the states of w
        # every `0 +` of the reference is there (the per-step kernels keep them, see jit_source)
        assert "hb_add(0.0, " in k or "= 0.0;" in k, i
    assert any("for (unsigned gi = 0;" in k for k in blocked.values())
    off = subprocess.run([os.sys.executable, "-c", "import sys; sys.path.insert(0, %r)\n"
                          "import hbayes_b200 as hb\nfrom hbayes_b200 import synth\n"
                          "net, observed = synth.config_c3()\n"
                          "pnet = hb.BayesianNetwork.from_tables(net.dims, net.scopes, net.values)\n"
                          "src = hb.JunctionTree(pnet, device=None).jit_source(sorted(observed))\n"
                          "print(src.count('__launch_bounds__(256, 1)'), src.count('hb_step_'))" % os.path.dirname(G.rstrip("/")).rsplit("/tests", 1)[0]],
                         capture_output=True, text=True, env=dict(os.environ, HB_JIT_BLOCK_U="0"))
    assert off.returncode == 0, off.stderr[-2000:]
    n_blocked_off, n_kernels_off = map(int, off.stdout.split())
    assert n_blocked_off == 0 and n_kernels_off < len(ks)  # the looped steps are only eligible with a block
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc")
    cu = tmp_path / "blocked.cu"
    cu.write_text(src.split('extern "C"', 1)[0] + "".join(blocked[i] for i in sorted(blocked)))
    r = subprocess.run([nvcc, "--fmad=false", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-Xptxas", "-v", str(cu), "-o", str(tmp_path / "blocked.cubin")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    spills = [int(x) for x in re.findall(r"(\d+) bytes spill stores", r.stderr)]
    assert len(spills) == len(blocked) and max(spills) <= 512, spills


def test_on_chip_kernel_leaves_out_the_zero_adds_only_for_nonnegative_potentials():
    """csrc/jit.cpp jit_zero_add_elidable: the reference starts every sum at 0 (`sum` = foldl (+) 0).  0 + x == x bit for bit
    unless x is -0.0 (or a signalling NaN), and no intermediate can be either when no potential is negative, -0.0 or NaN:
    then hb_tree leaves the adds out.  One -0.0 anywhere in a CPT and they are all back."""
    net = random_net(11, n=10, states=(2, 3), max_parents=3)
    values = [np.asarray(v, np.float64).copy() for v in net.values]
    zero_adds = lambda s: s.count("hb_add(zero, ") + s.count("hb_add(0.0, ") + s.count("__dadd_rn(0.0, ")
    plain = hb.JunctionTree(hb.BayesianNetwork.from_tables(net.dims, net.scopes, values), device=None).tree_source([1, 4])
    assert zero_adds(plain) == 0
    for bad in (-0.0, -0.25, float("nan")):
        v2 = [v.copy() for v in values]
        v2[3][0] = bad
        src = hb.JunctionTree(hb.BayesianNetwork.from_tables(net.dims, net.scopes, v2), device=None).tree_source([1, 4])
        assert zero_adds(src) > 20, bad
        assert src.count("hb_add(") > plain.count("hb_add(")
