"""CPU-only: the R `.Call` shim (r/src/frk_b200_shim.c) against the reference's registration.
R is not installed here; the shim is compiled against declarations of R's C API and linked against a mock of it."""
import json
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
REF_EXPORTS = Path("/root/reference/src/RcppExports.cpp")
REF_R_EXPORTS = Path("/root/reference/R/RcppExports.R")
GOLD = ROOT / "tests" / "golden" / "ref_callentries.json"


def reference_table():
    """name -> arity of the reference's R_CallMethodDef table (src/RcppExports.cpp:71-77) and the arities its R wrappers
    pass (R/RcppExports.R:4-18); parsed from the reference when it is present, else the committed copy of that parse."""
    if REF_EXPORTS.exists():
        txt = REF_EXPORTS.read_text()
        table = {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(_autoFRK_\w+)",\s*\(DL_FUNC\)\s*&\w+,\s*(\d+)\}', txt)}
        rtxt = REF_R_EXPORTS.read_text()
        rcalls = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
                  for m in re.finditer(r"\.Call\(`(_autoFRK_\w+)`((?:,\s*\w+)*)\)", rtxt)}
        got = {"call_entries": table, "r_wrapper_nargs": rcalls, "init": "R_init_autoFRK" in txt}
        if not GOLD.exists() or json.loads(GOLD.read_text()) != got:
            GOLD.write_text(json.dumps(got, indent=1))
        return got
    return json.loads(GOLD.read_text())


def test_shim_compiles_against_the_r_api_declarations():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Werror", f"-I{ROOT / 'tests' / 'r_stub'}", f"-I{ROOT / 'include'}",
                        str(ROOT / "r" / "src" / "frk_b200_shim.c")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_registration_table_matches_the_reference():
    from r_shim_util import MockR

    ref = reference_table()
    assert ref["init"] and len(ref["call_entries"]) == 4
    assert ref["call_entries"] == ref["r_wrapper_nargs"]          # the reference's own R wrappers agree with its table
    mine = MockR().routines()                                      # what R_init_autoFRK of the shim registers
    for name, nargs in ref["call_entries"].items():
        assert mine.get(name) == nargs, (name, nargs, mine.get(name))
    extra = set(mine) - set(ref["call_entries"])
    assert extra == {"_autoFRK_gram_stats", "_autoFRK_getHalf_stats", "_autoFRK_cMLEimat_stats", "_autoFRK_cMLEimat_sweep",
                     "_autoFRK_predict_mrts_basis", "_autoFRK_predict_gram", "_autoFRK_predict_frk",
                     "_autoFRK_masked_grams", "_autoFRK_krige_V"}
    # every .Call in the overlay's R code names a registered routine and passes its registered number of arguments
    rsrc = (ROOT / "r" / "R" / "overrides.R").read_text()
    calls = re.findall(r"\.Call\(`(_autoFRK_\w+)`", rsrc)
    assert calls and set(calls) <= set(mine)
    for m in re.finditer(r"\.Call\(`(_autoFRK_\w+)`,", rsrc):
        depth, i, nargs = 1, m.end(), 1
        while depth:
            ch = rsrc[i]
            depth += ch in "([" 
            depth -= ch in ")]"
            nargs += (ch == "," and depth == 1)
            i += 1
        assert nargs == mine[m.group(1)], (m.group(1), nargs, mine[m.group(1)])


def test_dotcall_enforces_names_and_arity_like_r():
    from r_shim_util import MockR

    R = MockR()
    with pytest.raises(RuntimeError, match="not available for .Call"):
        R.call("_autoFRK_nope", R.nil())
    with pytest.raises(RuntimeError, match="Incorrect number of arguments \\(2\\), expecting 3"):
        R.call("_autoFRK_mrtsrcpp", R.nil(), R.nil())
    # integer input is an error exactly as with Eigen::Map<MatrixXd>: raised before any device work
    with pytest.raises(RuntimeError, match="Wrong R type for mapped matrix"):
        R.call("_autoFRK_getASCeigens", R.imat([[1, 2], [2, 1]]))
    R.free()
