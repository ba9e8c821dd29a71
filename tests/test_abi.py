"""CPU tests of the boundary: the C-ABI library loads, exports every symbol that
include/cabanapic_b200.h declares, validates arguments, and refuses to run without a GPU
(no silent CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header="cabanapic_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cpic_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import cabanapic_b200 as m
    L = m.lib()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"
    assert sorted(m._lib.EXPORTED) == names, "python binding list and header disagree"
    assert L.cpic_abi_version() == 1
    # the multi-GPU layer (host C++ + NCCL inside the same library)
    mg = declared_functions("cabanapic_b200_mgpu.h")
    assert len(mg) >= 14
    for n in mg:
        assert hasattr(L, n), f"{n} declared in cabanapic_b200_mgpu.h but not exported"
    assert sorted(m._lib.EXPORTED_MGPU) == mg


def test_struct_layouts_match_header():
    import cabanapic_b200 as m
    assert ctypes.sizeof(m._lib.Consts) == 13 * 8
    assert ctypes.sizeof(m._lib.Params) == 10 * 4 + 8 + 8 * 4
    assert ctypes.sizeof(m._lib.PushStats) == 8 * 8


def test_argument_validation_without_gpu():
    """Validation happens before any CUDA call, so it is testable here."""
    import cabanapic_b200 as m
    for kw, code in ((dict(ng=2), -1), (dict(boundary=m.BOUNDARY_REFLECT, solver=m.SOLVER_ES_1D), -6), (dict(boundary=7), -1),
                     (dict(solver=m.SOLVER_ES_1D, ny=4), -1), (dict(real="f2"), None)):
        args = dict(nx=4, ny=1, nz=1, ng=1)
        args.update({k: v for k, v in kw.items() if k != "real"})
        if code is None:
            continue
        with pytest.raises(m.CpicError) as e:
            m.Context(**args)
        assert e.value.code == code


def test_no_cpu_fallback():
    """Without a CUDA device context creation must fail loudly (CPIC_E_CUDA), never compute."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cabanapic_b200 as m
    with pytest.raises(m.CpicError) as e:
        m.Context(4, 4, 4, 1, max_particles=10)
    assert e.value.code == -2
    assert "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure; the package must not reference it."""
    pkg = os.path.join(ROOT, "cabanapic_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("the oracle", "").replace("CPU oracle", ""), f


def test_benched_kernel_sass_matches_the_recorded_hash():
    """The SASS of the benched k_push3 instantiation in the in-tree library equals the hash recorded next to the profiles
    (profiles/k_push3_sass.md5, written by `tools/sass_hash.sh record`): what was measured is what ships, and a build with
    a developer flag left on (tools/variants/*.patch) cannot slip in unnoticed."""
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    want = open(os.path.join(root, "profiles", "k_push3_sass.md5")).read().strip()
    got = subprocess.run([os.path.join(root, "tools", "sass_hash.sh")], capture_output=True, text=True).stdout.strip()
    assert got == want, "k_push3 SASS changed: re-measure, then tools/sass_hash.sh record"


def test_variant_patches_apply_to_this_tree():
    """tools/variants/*.patch (kernel variants that were measured and not adopted, developer instrumentation) are kept
    applicable: the numbers in DESIGN.md that come from them stay reproducible."""
    import glob
    import shutil
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not shutil.which("git") or not os.path.isdir(os.path.join(root, ".git")):
        pytest.skip("not a git work tree")
    patches = sorted(glob.glob(os.path.join(root, "tools", "variants", "*.patch")))
    assert patches
    for p in patches:
        r = subprocess.run(["git", "apply", "--check", p], cwd=root, capture_output=True, text=True)
        assert r.returncode == 0, (os.path.basename(p), r.stderr[-500:])
