"""Documentation that names code must keep naming code that exists."""

import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*parts):
    with open(os.path.join(ROOT, *parts), encoding="utf-8") as f:
        return f.read()


def test_every_diagnostic_switch_is_documented_and_every_documented_switch_exists():
    src = ""
    for pat in ("pychebyshev_b200/csrc/*.cu", "pychebyshev_b200/csrc/*.cuh", "pychebyshev_b200/csrc/*.inc",
                "pychebyshev_b200/*.py", "tests/*.py"):
        for path in glob.glob(os.path.join(ROOT, pat)):
            if not path.endswith("test_docs_sync.py"):
                src += _read(os.path.relpath(path, ROOT))
    used = set(re.findall(r'(?:getenv|env_int|environ\.get|environ\[|setenv|delenv)\(?\s*"(PCB_[A-Z0-9_]+)"', src))
    used -= {"PCB_TT_CHAIN_PART"}
    section = _read("INTEGRATION.md").split("## 9. Diagnostic switches")[1]
    documented = set()
    for name in re.findall(r"PCB_[A-Z0-9_]+(?:\[[A-Z_|]+\])?", section):
        m = re.match(r"(PCB_[A-Z0-9_]+)\[([A-Z_|]+)\]", name)
        if m:   # PCB_TT_QPT[_VALUE|_FD] -> three names
            documented |= {m.group(1)} | {m.group(1) + suffix for suffix in m.group(2).split("|")}
        else:
            documented.add(name)
    assert used <= documented, f"switches read by the code but missing from INTEGRATION.md: {sorted(used - documented)}"
    assert documented <= used, f"INTEGRATION.md documents switches nothing reads: {sorted(documented - used)}"


def test_profiles_index_names_existing_files():
    index = _read("profiles", "README.md")
    for name in re.findall(r"`(r[12]_[A-Za-z0-9_{},.]+|traffic\.json)`", index):
        m = re.match(r"(.*)\{(.*)\}(.*)", name)
        for cand in ([m.group(1) + x + m.group(3) for x in m.group(2).split(",")] if m else [name]):
            assert os.path.exists(os.path.join(ROOT, "profiles", cand)), f"profiles/README.md names {cand}"


def test_header_cites_reference_lines_for_every_entry_point_family():
    header = _read("include", "pcb_b200.h")
    for anchor in ("barycentric.py", "tensor_train.py", "spline.py", "slider.py", "_binary.py", "_extrude_slice.py"):
        assert anchor in header, f"include/pcb_b200.h no longer cites {anchor}"
