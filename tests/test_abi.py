"""The C-ABI boundary (include/nefnet_b200.h): the shared library loads without a GPU and exports every
symbol the header declares; the ctypes mirror covers the same set; the parameter table matches the
reference's state_dict contract.  No compute call is made here."""
import ctypes
import os
import re

import pytest

from oracle import nefnet_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nefnet_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nef_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from network import _native as N
    if not os.path.exists(N.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return N.load()


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("nef_init", "nef_forward", "nef_backward", "nef_gen_ecg", "nef_gconv_fwd", "nef_gconv_wgrad",
                 "nef_loss_fwd", "nef_loss_bwd", "nef_sgd_step", "nef_plan_create", "nef_plan_bind"):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    raw = ctypes.CDLL(lib._name)
    for name in _declared():
        assert hasattr(raw, name), "libnefnet_b200.so does not export %s" % name


def test_ctypes_mirror_covers_the_header():
    from network import _native as N
    assert sorted(N.SIGNATURES) == _declared()


def test_struct_sizes_match_the_compiled_layout(lib):
    """ctypes mirrors of the descriptor structs against sizeof() as the library was compiled."""
    from network import _native as N
    for which, cls in enumerate((N.NefConvTerm, N.NefConvDesc, N.NefWgradDesc, N.NefForwardArgs, N.NefBackwardArgs)):
        assert ctypes.sizeof(cls) == lib.nef_struct_size(which), cls.__name__


@pytest.mark.parametrize("G", [1, 3, 12])
def test_parameter_table_is_the_reference_state_dict(lib, G):
    shapes = O.param_shapes(G)
    n = lib.nef_param_count(G)
    assert n == len(shapes)
    for i, (name, shape) in enumerate(shapes.items()):
        assert lib.nef_param_name(G, i).decode() == name
        numel = 1
        for d in shape:
            numel *= d
        assert lib.nef_param_numel(G, i) == numel


def test_module_state_dict_keys_and_shapes():
    """The nn.Module mirror can be constructed (not run) on CPU; its state_dict is the reference's."""
    import network
    m = network.Model_nefnet(theta_encoder_len=1, lead_num=3)
    sd = m.state_dict()
    shapes = O.param_shapes(3)
    assert list(sd.keys()) == list(shapes.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(shapes[k]), k
    with pytest.raises(ValueError):
        network.Model_nefnet(theta_encoder_len=2, lead_num=3)


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from network import _native as N
    with pytest.raises(RuntimeError):
        N.init(0)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing in the product package, the C sources or the tools may import it, and
    bench.py only inside cpu_baseline() (the reported CPU leg)."""
    import ast
    import glob
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "electrocardio-panorama_b200")
    files = glob.glob(os.path.join(pkg, "**", "*.py"), recursive=True) + glob.glob(os.path.join(root, "tools", "*.py"))
    assert files
    for f in files:
        tree = ast.parse(open(f).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            assert not any(m.split(".")[0] == "oracle" for m in mods), f
    tree = ast.parse(open(os.path.join(root, "bench.py")).read())
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        for node in ast.walk(fn):
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle":
                assert fn.name == "cpu_baseline", fn.name
    for node in tree.body:   # no module-level import either
        assert not (isinstance(node, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(node))


def test_plan_sizing_is_host_only_and_fits_the_part(lib):
    """nef_plan_create / nef_plan_workspace_bytes make no CUDA call: the caller-owned workspace of every BASELINE
    configuration is known before a device exists, and fits the 180 GB of HBM3e (70.1 GB at C2)."""
    import ctypes as C
    sizes = {}
    for name, (B, G, L, V) in {"C2": (256, 12, 5000, 0), "C4": (64, 12, 20000, 0), "C5": (64, 12, 5000, 24),
                               "C5_one_gpu": (512, 12, 5000, 24), "tiny": (1, 1, 16, 0)}.items():
        h = C.c_void_p()
        assert lib.nef_plan_create(B, G, L, V, C.byref(h)) == 0, lib.nef_last_error()
        sizes[name] = lib.nef_plan_workspace_bytes(h)
        lib.nef_plan_destroy(h)
    assert 50.0 < sizes["C2"] / 1e9 < 75.0, sizes["C2"]   # 70.1 GB with the fp16 operand / gradient copies of every block
    assert all(v < 180e9 * 0.9 for v in sizes.values()) and sizes["tiny"] < 64e6
    assert sizes["C5"] < sizes["C2"] / 3      # nothing is saved for backward per view: the 24 views reuse one decoder slot
    for bad in ((0, 1, 16, 0), (1, 0, 16, 0), (1, 1, 18, 0), (1, 1, 8, 0)):
        h = C.c_void_p()
        assert lib.nef_plan_create(*bad, C.byref(h)) != 0
        assert b"nef_plan_create" in lib.nef_last_error()


def test_host_only_helpers(lib):
    """Layout arithmetic and switch validation that need no device (CBL4 rows = B * (L + 6))."""
    assert lib.nef_cbl4_rows(256, 1250) == 256 * (1250 + 6)
    assert lib.nef_cbl4_floats(128, 2, 40) == ((128 // 4) * 2 * 46 + 528) * 4      # + tail guard rows
    assert lib.nef_prepare_scratch_bytes(256) > 0
    assert lib.nef_set_dec1_terms(0) != 0 and b"nef_set_dec1_terms" in lib.nef_last_error()
    assert lib.nef_set_dec1_terms(4) != 0
    assert lib.nef_set_dec1_terms(3) == 0                                           # the default stays selected
    assert lib.nef_param_numel(12, 10**6) == -1 and lib.nef_param_name(12, -1) == b""
    assert lib.nef_struct_size(99) == 0
