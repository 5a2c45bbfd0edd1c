"""CPU checks of the C++ host front-end against the oracle: parameter parsing, matrix assembly, projector
tables and bar right-hand-side data are identical (no device involved)."""
import os
import tempfile

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import ROOT, bar_data, default_params, make_problem, mo, pkg


def _load_csr(d, name):
    shape = np.fromfile(os.path.join(d, name + ".shape"), dtype=np.int64)
    rp = np.fromfile(os.path.join(d, name + ".rp"), dtype=np.int64)
    col = np.fromfile(os.path.join(d, name + ".col"), dtype=np.int32)
    val = np.fromfile(os.path.join(d, name + ".val"), dtype=np.float64)
    return sp.csr_matrix((val, col, rp), shape=tuple(shape))


def _maxdiff(a, b):
    d = (a - b).tocoo()
    return 0.0 if d.nnz == 0 else float(np.abs(d.data).max())


def _check_setup_against_oracle(pfile, prob, k, cycle):
    lib = pkg().load_host()
    for sd in prob.subs:
        with tempfile.TemporaryDirectory() as d:
            rc = lib.mmmfe_host_dump_subdomain(pfile.encode(), cycle, sd.r, k, d.encode())
            assert rc == 0, lib.mmmfe_host_last_error()
            assert list(np.fromfile(os.path.join(d, "neighbors"), dtype=np.int32)) == sd.neighbors
            K = _load_csr(d, "matrix")
            assert K.shape == sd.matrix.shape
            assert _maxdiff(K, sd.matrix.tocsr()) < 1e-14 * abs(sd.matrix).max()
            for side in range(4):
                if sd.neighbors[side] < 0:
                    continue
                f2c, c2f = _load_csr(d, f"f2c{side}"), _load_csr(d, f"c2f{side}")
                assert _maxdiff(f2c, sd.f2c[side]) < 1e-13 * abs(sd.f2c[side]).max()
                assert _maxdiff(c2f, sd.c2f[side]) < 1e-13 * abs(sd.c2f[side]).max()
            p0, src, fr_dofs, fr_vals = bar_data(prob, sd)
            np.testing.assert_allclose(np.fromfile(os.path.join(d, "p0")), p0, atol=1e-15)
            np.testing.assert_allclose(np.fromfile(os.path.join(d, "src")).reshape(src.shape), src, rtol=1e-12, atol=1e-16)
            got_dofs = np.fromfile(os.path.join(d, "fr_dofs"), dtype=np.int32)
            assert len(set(got_dofs)) == len(got_dofs)  # the device adds these entries without atomics
            assert sorted(got_dofs) == sorted(fr_dofs)
            got = np.fromfile(os.path.join(d, "fr_vals")).reshape(fr_vals.shape)
            np.testing.assert_allclose(got[:, np.argsort(got_dofs)], fr_vals[:, np.argsort(fr_dofs)], rtol=1e-12, atol=1e-16)


@pytest.mark.parametrize("k,cycle", [(1, 0), (1, 1), (2, 0), (2, 2)])
def test_front_end_setup_matches_oracle(k, cycle):
    prm = default_params(k, cycle + 1)
    _check_setup_against_oracle(os.path.join(ROOT, "parameter.txt"), make_problem(prm, cycle), k, cycle)


@pytest.mark.parametrize("cycle", [0, 1, 2, 3])
def test_front_end_dealii91_tie_rule_matches_oracle(monkeypatch, cycle):
    """MMMFE_TIE_RULE=dealii91: the C++ restatement of deal.II's point location picks the same cell as the oracle's
    (oracle/dealii_point_location.py, pinned to report.pdf Table 4 rows 1, 3, 5) at every tie point -- the tables
    are compared entry for entry -- and differs from the default rule from cycle 1 on."""
    monkeypatch.setenv("MMMFE_TIE_RULE", "dealii91")
    prm = default_params(2, cycle + 1)
    mesh = mo.refined_meshes(prm, 4)[cycle]
    prob = mo.Problem(prm, 4, mesh[:-1], mesh[-1], keep_solution=True, tie_rules="dealii91")
    prob.setup()
    _check_setup_against_oracle(os.path.join(ROOT, "parameter.txt"), prob, 2, cycle)
    ref = make_problem(prm, cycle)
    moved = sum(_maxdiff(a.f2c[s], b.f2c[s]) > 0 for a, b in zip(prob.subs, ref.subs) for s in a.f2c)
    assert (moved > 0) == (cycle >= 1)


def test_front_end_dealii91_tie_rule_on_an_8x8_layout(monkeypatch, tmp_path):
    """The same on the scaled-down configs[3] (8 x 8 checkerboard, non-cubic meshes, every side type)."""
    monkeypatch.setenv("MMMFE_TIE_RULE", "dealii91")
    mesh = [[8, 8, 16] if ((r % 8) + (r // 8)) % 2 == 0 else [4, 4, 8] for r in range(64)]
    p = pkg().write_parameter_file(str(tmp_path / "parameter.txt"), mesh, [2, 2, 4], mortar_degree=2)
    prm = mo.parse_parameter_file(p)
    prob = mo.Problem(prm, 64, mesh, [2, 2, 4], tie_rules="dealii91")
    lib = pkg().load_host()
    for r in (0, 7, 9, 27, 36, 63):
        sd = prob.subs[r]
        with tempfile.TemporaryDirectory() as d:
            assert lib.mmmfe_host_dump_subdomain(p.encode(), 0, r, 2, d.encode()) == 0, lib.mmmfe_host_last_error()
            for side in range(4):
                if sd.neighbors[side] < 0:
                    continue
                f2c, _ = mo.projector_tables(sd, side, [2, 2, 4], 2)
                assert _maxdiff(_load_csr(d, f"f2c{side}"), f2c) < 1e-13 * abs(f2c).max(), (r, side)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_front_end_dealii91_tie_rule_on_random_layouts(monkeypatch, tmp_path, seed):
    """Random decompositions (also 3 x 2 and 1 x 3: subdomain widths of 1/3), non-cubic meshes, mortar meshes with
    ties in space and time, final times 0.3 / 0.5 / 1: every f2c table of the C++ restatement equals the oracle's."""
    monkeypatch.setenv("MMMFE_TIE_RULE", "dealii91")
    rng = np.random.default_rng(seed)
    n0, n1 = [(3, 2), (2, 3), (1, 3), (4, 2)][seed]
    mort = [int(rng.integers(1, 4)) for _ in range(3)]
    mesh = [[mort[0] * 2 * int(rng.integers(1, 4)), mort[1] * 2 * int(rng.integers(1, 4)), mort[2] * 2 * int(rng.integers(1, 3))]
            for _ in range(n0 * n1)]
    p = pkg().write_parameter_file(str(tmp_path / "parameter.txt"), mesh, mort, mortar_degree=2,
                                   final_time=float(rng.choice([0.5, 1.0, 0.3])))
    prob = mo.Problem(mo.parse_parameter_file(p), n0 * n1, mesh, mort, tie_rules="dealii91")
    lib = pkg().load_host()
    sides = moved = 0
    for sd in prob.subs:
        with tempfile.TemporaryDirectory() as d:
            assert lib.mmmfe_host_dump_subdomain(p.encode(), 0, sd.r, 2, d.encode()) == 0, lib.mmmfe_host_last_error()
            for side in range(4):
                if sd.neighbors[side] < 0:
                    continue
                f2c, _ = mo.projector_tables(sd, side, mort, 2)
                assert _maxdiff(_load_csr(d, f"f2c{side}"), f2c) < 1e-13 * abs(f2c).max(), (sd.r, side)
                sides += 1
                moved += sum(1 for v in sd.tie_points.get(side, {}).values() if any(v))
    assert sides >= 4 and moved > 0


def test_front_end_rejects_an_unknown_tie_rule(monkeypatch):
    monkeypatch.setenv("MMMFE_TIE_RULE", "nearest")
    lib = pkg().load_host()
    with tempfile.TemporaryDirectory() as d:
        assert lib.mmmfe_host_dump_subdomain(os.path.join(ROOT, "parameter.txt").encode(), 0, 0, 2, d.encode()) != 0
        assert b"MMMFE_TIE_RULE" in lib.mmmfe_host_last_error()


@pytest.mark.parametrize("bc", [("N 0.5", "D 1.0", "D 2.0", "N -0.25"), ("D 1.0", "N 0.0", "N 1.5", "D 0.0")])
def test_front_end_neumann_sides_match_oracle(tmp_path, bc):
    """is_manufact_soln = 0 with essential (Neumann) outer sides: constrained matrix, system_rhs_bar_bc and the
    constrained rows (src/darcy_vtdd.cc:205-275, 479, 863-866); nx = 1 on one subdomain puts the row that receives
    the boundary term on an interface."""
    mesh, mortar = [[3, 3, 3], [1, 2, 2], [4, 4, 4], [3, 2, 3]], [1, 1, 1]
    p = pkg().write_parameter_file(str(tmp_path / "parameter.txt"), mesh, mortar, is_manufact=0, bc=bc)
    prm = mo.parse_parameter_file(p)
    assert not prm.is_manufact and prm.bc_type == [b.split()[0] for b in bc]
    prob = make_problem(prm, 0)
    assert sum(len(sd.con_dofs) for sd in prob.subs) > 0
    _check_setup_against_oracle(p, prob, -1, 0)


def test_parameter_file_round_trip(tmp_path):
    p = pkg().write_parameter_file(str(tmp_path / "parameter.txt"), [[8, 8, 4], [4, 4, 2]], [2, 2, 1], mortar_degree=2,
                                   num_refinement=3)
    prm = mo.parse_parameter_file(p)
    assert prm.mesh == [[8, 8, 4], [4, 4, 2], [2, 2, 1]] and prm.mortar_degree == 2 and prm.num_refinement == 3
    lib = pkg().load_host()
    assert lib.mmmfe_host_dump_subdomain(p.encode(), 0, 5, -1, str(tmp_path).encode()) == -1
    assert b"out of range" in lib.mmmfe_host_last_error()


def test_front_end_refuses_to_run_without_a_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg().MmmfeError, match="no CPU fallback"):
        pkg().host_run(os.path.join(ROOT, "parameter.txt"), verbose=0, write_output=0)


def test_reference_data_headers_compile_against_the_shim():
    """inc/data*.h of the reference must build unmodified against host/shim (drop-in surface)."""
    import glob
    import subprocess
    files = sorted(glob.glob("/root/reference/inc/data*.h"))
    if not files:
        pytest.skip("/root/reference is not present on this machine")
    shim = os.path.join(ROOT, "mmmfe-st-dd_b200", "host", "shim")
    for f in files:
        src = f'#include "{f}"\nint main(){{ vt_darcy::ExactSolution<2> e; (void)e; return 0; }}\n'
        r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", shim, "-x", "c++", "-"], input=src, text=True,
                           capture_output=True)
        assert r.returncode == 0, (f, r.stderr[:500])


def test_reference_data_headers_compile_for_the_device():
    """The same unmodified headers must also build for sm_100a through host/data_device.cu (bar data, error sums and
    output patches are then evaluated by CUDA kernels).  Three of the ten here to keep the CPU suite short; all ten
    were compiled when the path was written."""
    import shutil
    import subprocess
    import tempfile
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    files = [f"/root/reference/inc/{n}" for n in ("data.h", "data_sinx_siny.h", "data_const.h")]
    if not all(os.path.exists(f) for f in files):
        pytest.skip("/root/reference is not present on this machine")
    host = os.path.join(ROOT, "mmmfe-st-dd_b200", "host")
    with tempfile.TemporaryDirectory() as tmp:
        procs = [subprocess.Popen([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-fmad=false",
                                   "-diag-suppress", "20012,20014,940,20011,20015,177", "-I", os.path.join(host, "shim"),
                                   "-I", host, "-I", os.path.join(ROOT, "include"), f'-DMMMFE_DATA_HEADER="{f}"', "-c",
                                   os.path.join(host, "data_device.cu"), "-o", os.path.join(tmp, f"d{i}.o")],
                                  stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for i, f in enumerate(files)]
        for f, p in zip(files, procs):
            out = p.communicate()[0]
            assert p.returncode == 0, (f, out[:800])


def test_a_data_header_that_cannot_run_on_the_device_keeps_the_host_loops(tmp_path):
    """build.py's fallback: a header whose overriders call a plain host helper does not compile for sm_100a; the same
    translation unit with MMMFE_NO_DEVICE_DATA (stubs, `enabled()` = false) does, and the front-end then evaluates the
    data classes on the host as before."""
    import shutil
    import subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    host = os.path.join(ROOT, "mmmfe-st-dd_b200", "host")
    default = open(os.path.join(host, "data_default.h")).read()
    # a host-only helper (no MMMFE_HD) inside the right-hand side
    bad = default.replace("MMMFE_HD inline double pressure(", "inline double host_only_scale() { return 1.0; }\nMMMFE_HD inline double pressure(")
    bad = bad.replace("return 8 * std::cos(8 * t) * shape", "return host_only_scale() * 8 * std::cos(8 * t) * shape")
    assert bad != default
    hdr = tmp_path / "data_host_only.h"
    hdr.write_text(bad.replace("MMMFE_DATA_DEFAULT_H", "MMMFE_DATA_HOST_ONLY_H"))
    base = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-fmad=false", "-diag-suppress",
            "20012,20014,940,20011,20015,177", "-I", os.path.join(host, "shim"), "-I", host, "-I", os.path.join(ROOT, "include"),
            f'-DMMMFE_DATA_HEADER="{hdr}"', "-c", os.path.join(host, "data_device.cu")]
    dev = subprocess.run(base + ["-o", str(tmp_path / "dev.o")], capture_output=True, text=True)
    assert dev.returncode != 0 and "host_only_scale" in (dev.stdout + dev.stderr)
    stub = subprocess.run(base + ["-DMMMFE_NO_DEVICE_DATA", "-o", str(tmp_path / "stub.o")], capture_output=True, text=True)
    assert stub.returncode == 0, stub.stderr[:500]
