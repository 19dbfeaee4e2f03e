#!/usr/bin/env python3
"""Regenerate the DATA fixtures this repo needs from the (read-only) reference checkout.

Only data is extracted -- no reference source code is copied:
  * the 200-entry Boys F_10 lookup table (numbers in /root/reference/boys/src/lut.rs:1-202;
    they carry the reference's own quadrature noise, e.g. entry 0 is 4.7619047619031046e-2
    and not 1/21, so they cannot be regenerated from first principles),
  * trimmed MolSSI-BSE basis-set JSON (same schema the reference's serde structs read,
    basis_set/src/lib.rs:15-76) restricted to the 8 elements `AtomType::from_symbol`
    knows (basis_set/src/periodic_table.rs:284-296),
  * the sample XYZ geometries (chemfiles/molecules/*.xyz).

Run in the build container (where /root/reference exists):  python tools/extract_data.py
The outputs are committed; nothing at run time reads /root/reference.
"""
import json, os, re, sys

REF = os.environ.get("FOCK_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEEP_Z = ["1", "2", "5", "6", "7", "8", "9", "26"]
BASIS_FILES = ["STO-3G.json", "STO-6G.json", "3-21G.json", "6-31G.json", "6-31G_st.json", "6-31G_st_st.json", "6-311G.json"]
MOLECULES = ["water.xyz", "benzene.xyz", "ethane.xyz", "ethene.xyz", "ethyne.xyz", "ethanol.xyz", "cyclohexane.xyz"]


def extract_lut():
    src = open(os.path.join(REF, "boys/src/lut.rs")).read()
    vals = re.findall(r"^\s*([0-9][0-9.eE+\-]*),\s*$", src, flags=re.M)
    assert len(vals) == 200, len(vals)
    body = ",\n".join("  " + v for v in vals)
    header = (
        "/* Boys F_10(k*0.15), k=0..199: DATA extracted by tools/extract_data.py from the reference's\n"
        "   boys/src/lut.rs:1-202 (the reference's own quadrature output; reproduced digit for digit). */\n"
    )
    for out in ("oracle/boys_lut.inc", "fock_rs_b200/csrc/boys_lut.inc"):
        with open(os.path.join(ROOT, out), "w") as f:
            f.write(header + body + "\n")


def extract_basis():
    for name in BASIS_FILES:
        d = json.load(open(os.path.join(REF, "basis_set/src/data", name)))
        d["elements"] = {z: e for z, e in d["elements"].items() if z in KEEP_Z}
        for e in d["elements"].values():
            e["references"] = []
        d.pop("auxiliaries", None)
        out = os.path.join(ROOT, "fock_rs_b200/data/basis", name)
        with open(out, "w") as f:
            json.dump(d, f, separators=(",", ":"))
            f.write("\n")


def extract_molecules():
    for name in MOLECULES:
        src = open(os.path.join(REF, "chemfiles/molecules", name)).read()
        with open(os.path.join(ROOT, "fock_rs_b200/data/molecules", name), "w") as f:
            f.write(src)


if __name__ == "__main__":
    extract_lut()
    extract_basis()
    extract_molecules()
    print("ok")
