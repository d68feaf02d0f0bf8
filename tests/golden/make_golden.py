"""Generate tests/golden/*.json from the reference tree.  Run in the build container only
(needs /root/reference); the JSON files are committed so the GPU box never reads the reference.

  python tests/golden/make_golden.py

ragusa16.json : the only fixture the reference ships (data/real/A_matrix.mtx, B_matrix.mtx) as COO
                triples, plus the expected bmSparse structures.  Expected values are the ones
                SURVEY.md Appendix E derived independently (scipy cross-check); this script
                re-derives them with scipy and asserts the oracle agrees before writing.
cusp_*.json   : outputs of the reference's own cusp host CSR kernels (oracle/_ref/libcusp_ref.so,
                compiled from /root/reference by oracle/Makefile) on small gallery inputs.
"""
import json
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import oracle as O  # noqa: E402

REF = "/root/reference"


def hexs(a):
    return [f"{int(x):016x}" for x in a]


def ragusa():
    nr, nc, r, c, v = O.read_mtx(f"{REF}/data/real/A_matrix.mtx")
    _, _, rb, cb, vb = O.read_mtx(f"{REF}/data/real/B_matrix.mtx")
    A = O.coo_to_bmsp(nr, nc, r, c, v)
    Bt = O.coo_to_bmsp(nr, nc, rb, cb, vb, transposed=True)
    Cm = O.spgemm(A, Bt)
    As = sp.coo_matrix((v, (r, c)), shape=(nr, nc)).tocsr()
    Bs = sp.coo_matrix((vb, (rb, cb)), shape=(nr, nc)).tocsr()
    Cs = (As @ Bs).tocsr()
    # structural product (boolean) must equal the value product's pattern on this fixture
    Pat = (abs(As) @ abs(Bs)).tocsr()
    rr, cc, vv = O.bmsp_to_coo(Cm)
    Co = sp.coo_matrix((vv, (rr, cc)), shape=(nr, nc)).tocsr()
    assert Co.nnz == Pat.nnz == 255 and abs(Co - Cs).max() == 0
    y = As @ np.ones(nc)
    assert np.array_equal(O.spmv(A, np.ones(nc, np.float32)), y)
    out = dict(
        source="GonzaBerger/bmSparse-SPGEMM-SPMV data/real/{A,B}_matrix.mtx (Pajek/Ragusa16, 24x24, 81 nnz)",
        num_rows=nr, num_cols=nc,
        A=dict(rows=r.tolist(), cols=c.tolist(), vals=v.tolist()),
        B=dict(rows=rb.tolist(), cols=cb.tolist(), vals=vb.tolist()),
        A_keys=hexs(A.keys), A_bmps=hexs(A.bmps), A_offsets=[int(x) for x in A.offsets],
        Bt_keys=hexs(Bt.keys), Bt_bmps=hexs(Bt.bmps), Bt_offsets=[int(x) for x in Bt.offsets],
        C_keys=hexs(Cm.keys), C_bmps=hexs(Cm.bmps), C_offsets=[int(x) for x in Cm.offsets],
        C_coo=dict(rows=Cs.tocoo().row.tolist(), cols=Cs.tocoo().col.tolist(), vals=Cs.tocoo().data.tolist()),
        spmv_ones=y.tolist(),
        # SURVEY.md Appendix E, typed in from the survey (independent derivation)
        survey_A_bmps="0800080028000609 00000100be002470 0400010001000011 0a003b00014a0008 8030603030200081 "
                      "0c04300404040000 0000080d00300008 10200080003c14e0 0000201000140004".split(),
        survey_Bt_bmps="00000800a9020201 08010b09080a0820 0000000100800029 00042020a500a428 81207c5800000001 "
                       "0000202080dc0000 0000040431100010 1101458604060000 0000201400050000".split(),
        survey_C_bmps="380028007b007f3f be00ff00ff0064fe 150005003d003435 3e3b3f3b3b3f002a be7cff7c7cfe00bf "
                      "1d3435343535000d 003b282f007f4a3b 3060befe00fd30fe 0430211d0035043d".split(),
        survey_A_offsets=[0, 8, 20, 25, 37, 49, 57, 64, 76],
        survey_C_offsets=[0, 24, 56, 73, 108, 151, 176, 203, 236, 255],
        survey_spmv_ones=[3, 0, 3, 0, 15, 0, 4, 7, 5, 4, 12, 7, 4, 8, 0, 4, 1, 1, 2, 5, 0, 14, 2, 8],
    )
    json.dump(out, open(os.path.join(HERE, "ragusa16.json"), "w"), indent=0)


def cusp():
    cases = {}
    for name, (m, n) in {"poisson5pt_4x6": (4, 6), "poisson5pt_8x3": (8, 3), "poisson5pt_10x10": (10, 10)}.items():
        rp, ci, v = O.poisson5pt(m, n)
        N = m * n
        x = (np.arange(N) % 10).astype(np.float32)
        y = O.ref_csr_spmv(N, N, rp, ci, v, x)
        crp, cci, cv = O.ref_csr_spgemm(N, N, rp, ci, v, N, N, rp, ci, v, omp=False)
        orp, oci, ov = O.ref_csr_spgemm(N, N, rp, ci, v, N, N, rp, ci, v, omp=True)
        # scipy agrees with the reference kernels (pattern + values)
        S = sp.csr_matrix((v, ci, rp), shape=(N, N))
        assert abs(sp.csr_matrix((cv, cci, crp), shape=(N, N)) - S @ S).max() == 0
        cases[name] = dict(m=m, n=n, rp=rp.tolist(), ci=ci.tolist(), v=v.tolist(), x=x.tolist(), y=y.tolist(),
                           seq=dict(rp=crp.tolist(), ci=cci.tolist(), v=cv.tolist()),
                           omp=dict(rp=orp.tolist(), ci=oci.tolist(), v=ov.tolist()))
    json.dump(dict(source="oracle/_ref/libcusp_ref.so = cusp/cusp/system/{detail/sequential,omp/detail}/multiply/csr_*.h "
                          "compiled from the reference tree", cases=cases),
              open(os.path.join(HERE, "cusp_host.json"), "w"), indent=0)


if __name__ == "__main__":
    O.build()
    ragusa()
    cusp()
    print("golden written")
