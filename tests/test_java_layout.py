"""The Java marshaller of the drop-in (integration/java/.../JmeSystemLayout.java) cannot be compiled here (no JDK), but
the byte offsets it hard-codes for `jme_system_t` can be checked against the C header: a probe compiled with gcc prints
offsetof / sizeof of every field of include/jme_b200.h, and every OFF_* / SIZEOF constant of the Java source must agree
- so the unbuildable source is at least layout-checked."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "integration", "java", "org", "jme", "forcefield", "mmff", "JmeSystemLayout.java")
HEADER = os.path.join(ROOT, "include", "jme_b200.h")


def _struct_fields():
    """Field names of jme_system_t in declaration order, parsed from the header."""
    text = open(HEADER).read()
    body = text[text.index("typedef struct {") + len("typedef struct {"):text.index("} jme_system_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for stmt in body.split(";"):
        m = re.search(r"(\w+)\s*$", stmt.strip())
        if m and stmt.strip():
            names.append(m.group(1))
    return names


def test_java_offsets_match_the_c_struct(tmp_path):
    fields = _struct_fields()
    assert len(fields) == 38 and fields[0] == "n_atoms" and fields[-1] == "tor_v3"
    src = tmp_path / "probe.c"
    lines = ['#include <stddef.h>', '#include <stdio.h>', f'#include "{HEADER}"', "int main(void) {"]
    lines += [f'  printf("{f} %zu\\n", offsetof(jme_system_t, {f}));' for f in fields]
    lines += ['  printf("SIZEOF %zu\\n", sizeof(jme_system_t));', "  return 0;", "}"]
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    c_off = dict(l.split() for l in out.strip().splitlines())
    java = open(JAVA).read()
    j_off = dict(re.findall(r"static final long OFF_(\w+) = (\d+);", java))
    assert list(j_off) == fields, "the Java constants list the fields in the header's order"
    for f in fields:
        assert int(j_off[f]) == int(c_off[f]), (f, j_off[f], c_off[f])
    size = re.search(r"static final long SIZEOF_jme_system_t = (\d+);", java)
    assert size and int(size.group(1)) == int(c_off["SIZEOF"])
    # the StructLayout names the same fields in the same order
    named = re.findall(r'\.withName\("(\w+)"\)', java[java.index("StructLayout LAYOUT"):java.index(".withName(\"jme_system_t\")")])
    assert named == fields


def test_java_shim_binds_only_exported_symbols():
    """Every downcall handle of the FFM component names a symbol that include/jme_b200.h declares and the library exports."""
    comp = open(os.path.join(os.path.dirname(JAVA), "GpuMMFF94Component.java")).read()
    used = set(re.findall(r'h\("(jme_\w+)"', comp))
    assert used >= {"jme_create", "jme_destroy", "jme_set_options", "jme_set_coords", "jme_set_coord1", "jme_energy",
                    "jme_energy_gradient", "jme_last_error"}
    header = open(HEADER).read()
    sys.path.insert(0, ROOT)
    from jmolecularenergy_b200 import _lib
    for name in used:
        assert re.search(rf"\b{name}\s*\(", header), name
        assert name in _lib.EXPORTS, name
