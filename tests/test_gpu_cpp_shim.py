"""GPU: the C++ shim (include/bmSpMatrix.h / reader.h / CSRMatrix.h) driving the library the way the reference's mains do."""
import os
import subprocess

import pytest

from tests.util import load_golden

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_mtx(path, n, rows, cols, vals):
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n%\n")
        f.write(f"{n} {n} {len(rows)}\n")
        for r, c, v in zip(rows, cols, vals):
            f.write(f"{r + 1} {c + 1} {v}\n")


def test_shim_runs_reference_style_main(tmp_path):
    exe = os.path.join(ROOT, "tools", "_build", "shim_smoke")
    if not os.path.exists(exe):
        pytest.skip("tools/_build/shim_smoke not built")
    g = load_golden("ragusa16.json")
    a, b = str(tmp_path / "A.mtx"), str(tmp_path / "B.mtx")
    _write_mtx(a, 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])
    _write_mtx(b, 24, g["B"]["rows"], g["B"]["cols"], g["B"]["vals"])
    out = subprocess.run([exe, a, b], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert "C blocks: 9" in lines and "C nnz: 255" in lines
    assert f"SpMV sum: {sum(g['spmv_ones']):.1f}" in lines
    assert "mmread: 24 24 81 blocks 9" in lines and "CSR C nnz: 255" in lines


def test_cli_drivers_print_the_reference_lines(tmp_path):
    """cli/bmsparse_spgemm_float and cli/bmsparse_spmv_float take the batch scripts' arguments (folder, names without ".mtx",
    segmented / tc_version / verbose or batched) and print the reference's summary lines (SPGEMM.cu:1278-1285, SPMV.cu:302-306)."""
    spgemm = os.path.join(ROOT, "cli", "_build", "bmsparse_spgemm_float")
    spmv = os.path.join(ROOT, "cli", "_build", "bmsparse_spmv_float")
    if not (os.path.exists(spgemm) and os.path.exists(spmv)):
        pytest.skip("cli/_build not built")
    g = load_golden("ragusa16.json")
    _write_mtx(str(tmp_path / "A_matrix.mtx"), 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])
    _write_mtx(str(tmp_path / "B_matrix.mtx"), 24, g["B"]["rows"], g["B"]["cols"], g["B"]["vals"])
    out = subprocess.run([spgemm, str(tmp_path), "A_matrix", "B_matrix", "0", "5", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert "C blocks: 9" in lines and "C nnz: 255" in lines
    assert any(l.startswith("bmSparse execution: ") and l.endswith(" μs") for l in lines)
    assert any(l.startswith("T_7 (numeric): ") for l in lines)
    assert f"A matrix: {tmp_path}/A_matrix" in lines
    for args in ([str(tmp_path), "A_matrix", "A_matrix", "0"], [str(tmp_path), "A_matrix"]):
        out = subprocess.run([spmv] + args, capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stdout + out.stderr
        lines = out.stdout.splitlines()
        assert any(l.startswith("bmSparse SpMV execution: ") and l.endswith(" μs") for l in lines)
        assert f"y checksum: {sum(g['spmv_ones']):g}" in lines
    out = subprocess.run([spgemm, str(tmp_path), "A_matrix"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 1 and "MatrixFolder" in out.stdout
    out = subprocess.run([spgemm, str(tmp_path), "A_matrix", "missing"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 2 and "error" in out.stderr


def test_reference_mains_run_against_the_shim(tmp_path):
    """tools/_build/ref_main_spgemm and ref_main_spmv are the reference's own main() bodies (src/bmSparse_SPGEMM.cu:1226-1288,
    src/bmSparse_SPMV.cu:232-312), cut out of the reference tree at build time and compiled UNCHANGED against include/
    (tests/cpp/ref_main_wrapper.cu): same constructor calls, same operator calls, same member accesses.  Here they run."""
    spgemm = os.path.join(ROOT, "tools", "_build", "ref_main_spgemm")
    spmv = os.path.join(ROOT, "tools", "_build", "ref_main_spmv")
    if not (os.path.exists(spgemm) and os.path.exists(spmv)):
        pytest.skip("tools/_build/ref_main_* not built (reference tree absent at build time)")
    g = load_golden("ragusa16.json")
    _write_mtx(str(tmp_path / "A_matrix.mtx"), 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])
    _write_mtx(str(tmp_path / "B_matrix.mtx"), 24, g["B"]["rows"], g["B"]["cols"], g["B"]["vals"])
    _write_mtx(str(tmp_path / "A_matrix"), 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])     # SPMV.cu:270 opens the path without ".mtx"
    # the reference reads argv[argc] when optional arguments are missing: always pass all of them
    out = subprocess.run([spgemm, str(tmp_path), "A_matrix", "B_matrix", "0", "5", "1"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert "C blocks: 9" in lines and "C nnz: 255" in lines
    assert any(l.startswith("bmSparse execution: ") for l in lines)
    out = subprocess.run([spmv, str(tmp_path), "A_matrix", "0"], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    assert any(l.startswith("bmSparse SpMV execution: ") for l in out.stdout.splitlines())


def test_thrust_typed_overloads(tmp_path):
    """the reference's Thrust-typed signatures: five-argument mmread_bmSparse (include/reader.h:14-15) and the adopting constructor
    (include/bmSpMatrix.h:33-34), which leaves the caller's vectors empty like the reference's swap does"""
    exe = os.path.join(ROOT, "tools", "_build", "thrust_shim")
    if not os.path.exists(exe):
        pytest.skip("tools/_build/thrust_shim not built")
    g = load_golden("ragusa16.json")
    a = str(tmp_path / "A.mtx")
    _write_mtx(a, 24, g["A"]["rows"], g["A"]["cols"], g["A"]["vals"])
    out = subprocess.run([exe, a], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.splitlines()
    assert "mmread5: 24 24 81 blocks 9 values 81" in lines
    assert "adopted: blocks 9 nnz 81 caller vectors now 0 0 0 0" in lines
    assert f"SpMV sum: {sum(g['spmv_ones']):.1f}" in lines
