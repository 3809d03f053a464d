"""CPU: the C-ABI library loads and exports every symbol include/b200voc.h declares; the host-side
mirror keeps the reference's module / state_dict layout; compute fails loudly without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b200voc.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200voc_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from b200voc import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b200voc.h but not exported"
    assert set(syms) == set(_lib.EXPORTS), set(syms) ^ set(_lib.EXPORTS)
    assert _lib.load().b200voc_version() >= 100


def test_product_library_has_no_experiment_exports():
    """experiment / trace entry points live in the development build only (include/b200voc_dev.h, `make dev`)"""
    from b200voc import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    text = open(os.path.join(ROOT, "include", "b200voc_dev.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    dev_syms = sorted(set(re.findall(r"\b(b200voc_[a-z0-9_]+)\s*\(", text)))
    assert set(dev_syms) == set(_lib.DEV_EXPORTS)
    for s in dev_syms:
        assert not hasattr(lib, s), f"{s} must not be exported by the product library"
    if os.path.exists(_lib.DEV_LIB_PATH):
        dev = ctypes.CDLL(_lib.DEV_LIB_PATH)
        for s in dev_syms + list(_lib.EXPORTS):
            assert hasattr(dev, s), s


def test_state_dict_layout_matches_oracle_and_reference():
    from b200voc import GANConfig, Generator
    from oracle import vocoder7_oracle as O
    torch.manual_seed(1234)
    gen = Generator(GANConfig())
    ora = O.make_generator(O.OracleConfig(), seed=1234)
    a, b = gen.state_dict(), ora.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        assert torch.equal(a[k], b[k]), k          # same construction order -> same seeded init
    # spot-check the pinned reference shapes
    assert a["band_split.0.weight"].shape == (512, 20, 7)
    assert a["upsample_blocks.0.0.weight"].shape == (512, 256, 16)    # ConvT layout [in, out, k]
    assert a["upsample_blocks.3.0.weight"].shape == (64, 32, 4)
    assert a["band_merge.weight"].shape == (1, 128, 7)
    assert "upsample_blocks.2.4.q.weight" in a and "upsample_blocks.1.4.q.weight" not in a


def test_kernel_plan_expects_exactly_the_state_dict():
    from b200voc import GANConfig, Generator, _lib
    lib = _lib.load()
    gen = Generator(GANConfig())
    cc = gen._c_config()
    h = ctypes.c_void_p()
    rc = lib.b200voc_gen_create(ctypes.byref(cc), ctypes.byref(h))
    if rc != 0:   # no CUDA driver in this container: create needs device memory -> loud failure
        msg = lib.b200voc_last_error_string()
        assert rc == _lib.ERR_CUDA and (b"cudaMalloc" in msg or b"stream" in msg), msg
        return
    n = lib.b200voc_gen_num_weights(h)
    names = {lib.b200voc_gen_weight_name(h, i).decode(): lib.b200voc_gen_weight_numel(h, i) for i in range(n)}
    sd = gen.state_dict()
    assert set(names) == set(sd)
    for k, v in sd.items():
        assert names[k] == v.numel(), k
    lib.b200voc_gen_destroy(h)


def test_gst_module_layout_matches_reference():
    from b200voc import GANConfig, GlobalStyleTokens, _lib
    from oracle import vocoder7_oracle as O
    torch.manual_seed(1234)
    gst = GlobalStyleTokens(GANConfig())
    sd, ref = gst.state_dict(), O.make_gst_state(seed=1234)
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert torch.equal(sd[k], ref[k]), k
    with pytest.raises(_lib.B200VocError):          # no CPU path
        gst(torch.zeros(1, 80, 4))
    with pytest.raises(ValueError):
        gst(torch.zeros(1, 80, 4), mel_layout="TBC")


def test_no_cpu_fallback():
    from b200voc import GANConfig, Generator, _lib
    gen = Generator(GANConfig(use_attention=False))
    with pytest.raises(_lib.B200VocError):
        gen(torch.zeros(1, 80, 4), torch.zeros(1, 4, 18), torch.zeros(1, 128), torch.zeros(1, 6))


def test_bad_config_rejected():
    from b200voc import GANConfig, Generator, _lib
    lib = _lib.load()
    gen = Generator(GANConfig())
    cc = gen._c_config()
    cc.hidden_dim = 100
    h = ctypes.c_void_p()
    assert lib.b200voc_gen_create(ctypes.byref(cc), ctypes.byref(h)) == _lib.ERR_BAD_ARG
    assert b"hidden_dim" in lib.b200voc_last_error_string()
    with pytest.raises(ValueError):
        Generator(GANConfig(precision="int4"))._c_config()


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "tts-core-remastered-1_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, os.path.join(dp, f)
