"""CPU-side checks: drop-in boundary (constructor, state_dict keys, init parity with the reference's
construction order), C-ABI library loads and exports every declared symbol, host helpers."""
import ctypes
import os
import re

import pytest
import torch

from oracle import vqa_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_keys_shapes_match_reference_contract():
    import dl_vqa_b200 as D
    m = D.VqaNet(O.DEFAULT_CFG, 15000)
    want = O.param_shapes(O.DEFAULT_CFG, 15000)
    got = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    assert got == want
    assert sum(p.numel() for p in m.parameters()) == 23793242
    for name in ("text", "image", "attention", "classifier"):          # utils/main_utils.py:33-36
        assert sum(p.numel() for p in getattr(m, name).parameters()) > 0


def test_same_seed_init_equals_reference_init(golden_full):
    import dl_vqa_b200 as D
    torch.manual_seed(golden_full["seed"])
    sd = D.VqaNet(golden_full["cfg"], golden_full["V"]).state_dict()
    for k, d in golden_full["weight_digest"].items():
        assert torch.equal(sd[k].flatten()[:8], d["head"]), k
        assert abs(float(sd[k].double().sum()) - d["sum"]) < 1e-6 * max(1.0, abs(d["sum"])), k


def test_library_loads_and_exports_every_declared_symbol():
    from dl_vqa_b200 import lib
    hdr = open(os.path.join(ROOT, "include", "vqa_b200.h")).read()
    declared = set(re.findall(r"\b(vqa_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    dll = ctypes.CDLL(lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(dll, name), f"{name} declared in include/vqa_b200.h but not exported"
    dll.vqa_abi_version.restype = ctypes.c_int
    assert dll.vqa_abi_version() == 1
    lib.load()
    bound = set(lib.PROTOTYPES) | set(lib._optional_prototypes()) | {"vqa_last_error_string", "vqa_abi_version", "vqa_launch_count"}
    assert declared <= bound, f"declared but not bound in lib.py: {declared - bound}"


def test_ctypes_prototypes_match_the_header_signatures():
    """Every entry of include/vqa_b200.h against its ctypes prototype in dl_vqa_b200/lib.py / lib_tc.py: same number of
    parameters, same kind each (pointer, int, int64_t, uint64_t, uint32_t, float, double).  A drifted prototype would
    otherwise only show up on the GPU as a wrong argument."""
    from dl_vqa_b200 import lib
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "vqa_b200.h")).read(), flags=re.S)
    decls = re.findall(r"\b[A-Za-z_][A-Za-z0-9_ \*]*?\b(vqa_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S)
    assert len(decls) >= 60
    lib.load()
    protos = dict(lib.PROTOTYPES)
    protos.update(lib._optional_prototypes())
    scalar = {"int": ctypes.c_int, "int64_t": ctypes.c_int64, "uint64_t": ctypes.c_uint64, "uint32_t": ctypes.c_uint32,
              "float": ctypes.c_float, "double": ctypes.c_double}

    def kind(param):
        param = param.strip()
        if param in ("", "void"):
            return None
        if "*" in param:
            return ctypes.c_void_p
        ctype = re.sub(r"\b[A-Za-z_][A-Za-z0-9_]*$", "", param).replace("const", "").strip()      # drop the parameter name
        assert ctype in scalar, f"unknown C type {ctype!r} in the header"
        return scalar[ctype]

    checked = 0
    for name, params in decls:
        if name in ("vqa_last_error_string", "vqa_abi_version", "vqa_launch_count"):    # bound by hand in lib.load()
            continue
        want = [k for k in (kind(x) for x in params.split(",")) if k is not None]
        got = list(protos[name])
        assert len(got) == len(want), f"{name}: header has {len(want)} parameters, the prototype {len(got)}"
        for i, (g, w) in enumerate(zip(got, want)):
            assert g is w, f"{name}: parameter {i} is {w.__name__} in the header, {g.__name__} in the prototype"
        checked += 1
    assert checked >= 60


def test_no_cpu_fallback():
    import dl_vqa_b200 as D
    from dl_vqa_b200 import lib
    m = D.VqaNet(O.cfg_with(O.DEFAULT_CFG, image_size=64), 50)
    with pytest.raises(lib.VqaLibraryError):
        m(torch.zeros(1, 3, 64, 64), torch.ones(1, 4, dtype=torch.long), torch.tensor([4]))
    with pytest.raises(lib.VqaLibraryError):
        D.soft_target_loss_and_score(torch.zeros(2, 10), torch.zeros(2, 3, dtype=torch.long),
                                     torch.zeros(2, 3, dtype=torch.long))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dl_vqa_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no CPU or PyTorch fallback", ""), fn


def test_learning_rate_schedule():
    import dl_vqa_b200 as D
    p = torch.nn.Parameter(torch.zeros(3))
    opt = torch.optim.Adam([p], lr=5e-4)
    for it in (0, 1, 50000, 123456):
        lr = D.update_learning_rate(opt, it, 5e-4)
        assert lr == O.learning_rate(5e-4, it) == opt.param_groups[0]["lr"]


def test_synthetic_batch_shapes():
    from dl_vqa_b200 import synth
    v, q, ai, av, al, idx, ql = synth.make_batch(8)
    assert v.shape == (8, 3, 224, 224) and v.dtype == torch.float32 and torch.equal(v.half().float(), v)
    assert q.shape == (8, 23) and q.dtype == torch.int64 and int(ql[0]) == 23
    for b in range(8):
        assert (q[b, : int(ql[b])] > 0).all() and (q[b, int(ql[b]):] == 0).all()
        n = int(al[b])
        assert (ai[b, :n] > 0).all() and (ai[b, n:] == 0).all() and (av[b, n:] == 0).all()
        assert (ai[b, :n].diff() > 0).all() and int(av[b].sum()) <= 10


def test_vector_dropout_threshold_drops_the_requested_fraction():
    """nn.Dropout(p) (reference models/model.py:84,185,194) on the vector mask scheme: a 16-bit random field is compared AS A
    BF16 BIT PATTERN against a threshold (one HSET2 per two elements).  Over all 65536 patterns the threshold must drop
    round(p * 65536) of them -- exactly for the config values, within 2^-10 everywhere (thresholds are kept away from
    subnormals) -- and NaN patterns must count as kept.  Host code of the library only; no GPU work."""
    import ctypes as C
    from dl_vqa_b200 import lib
    L = lib.load()
    patterns = torch.arange(65536, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float()
    pat, cnt = C.c_uint32(0), C.c_uint32(0)
    for p in [0.0, 0.1, 0.2, 0.3, 0.4, 0.45, 0.4844, 0.497, 0.499, 0.5, 0.501, 0.6, 0.75, 0.9, 0.99]:
        assert L.vqa_dropout_threshold_pattern(C.c_float(p), C.byref(pat), C.byref(cnt)) == 0
        if p == 0.0:
            assert pat.value == 0
            continue
        assert cnt.value == int(p * 65536 + 0.5)
        t = torch.tensor([pat.value], dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float()
        assert bool(torch.isfinite(t).all()) and (float(t.abs()) == 0.0 or float(t.abs()) >= 2.0 ** -126)   # zero or normal
        dropped = int((patterns < t).sum())           # ordered compare: NaN patterns are never dropped
        tol = 0 if p in (0.1, 0.2, 0.3, 0.4, 0.6, 0.75, 0.9) else 64
        assert abs(dropped - cnt.value) <= tol, (p, dropped, cnt.value)
