"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol that
include/ucd_b200.h declares; host-side argument validation works without a GPU; the product refuses
CPU tensors instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    from ucd_b200 import _lib, build
    build.build()
    build.build(debug=True)
    return _lib


def declared_functions(header="ucd_b200.h"):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ucd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound(L):
    names = declared_functions()
    assert len(names) >= 20
    h = ctypes.CDLL(L.LIB_PATH)
    for n in names:
        assert hasattr(h, n), "libucd_b200.so does not export %s" % n
        assert n in L.EXPORTED, "ucd_b200/_lib.py has no ctypes prototype for %s" % n
    assert sorted(L.EXPORTED) == names, "ctypes prototypes and header disagree"


def test_debug_entry_points_only_in_the_debug_library(L):
    """Tracing, tuning knobs and the tcgen05 probes (include/ucd_b200_debug.h) are compiled into
    libucd_b200_debug.so only: the product library exports none of them, reads no environment variable and keeps
    no global debug state."""
    dbg = declared_functions("ucd_b200_debug.h")
    assert sorted(L.DEBUG_EXPORTED) == dbg and len(dbg) >= 5
    prod, debug = ctypes.CDLL(L.LIB_PATH), ctypes.CDLL(L.DEBUG_LIB_PATH)
    for n in dbg:
        assert not hasattr(prod, n), "product library exports the debug entry point %s" % n
        assert hasattr(debug, n), n
    for n in declared_functions():
        assert hasattr(debug, n), n
    blob = open(L.LIB_PATH, "rb").read()
    assert b"UCD_SPLITS" not in blob and b"UCD_UP_GY" not in blob   # (the static CUDA runtime has its own getenv)
    assert b"UCD_SPLITS1" in open(L.DEBUG_LIB_PATH, "rb").read()


def test_version_and_sizes(L):
    lib = L.lib()
    assert lib.ucd_version() >= 100
    assert lib.ucd_reduce_scratch_floats() > 0
    assert lib.ucd_con_max_tiles(3072) == 49
    assert lib.ucd_con_prob_kpad(16) == 16 and lib.ucd_con_prob_kpad(14) == 16 and lib.ucd_con_prob_kpad(17) == 32
    assert lib.ucd_con_workspace_bytes(24, 49, 0, 0) > 24 * 128 * 256 * 4 * 2
    assert lib.ucd_con_workspace_bytes(0, 49, 0, 0) == 0
    # a two-part run (local chunk + 7 remote chunks) needs partial slots for both launches
    assert lib.ucd_con_workspace_bytes(24, 8 * 49, 0, 49) > lib.ucd_con_workspace_bytes(24, 8 * 49, 0, 0)


def test_argument_validation_reports_errors(L):
    lib = L.lib()
    rc = lib.ucd_unce_fwd(None, None, None, None, None, None, None, 1, 1, 1, 1, 255, None)
    assert rc == -1 and b"null pointer" in lib.ucd_last_error()
    with pytest.raises(RuntimeError, match="null pointer"):
        L.check(rc, "unce_fwd")
    err = ctypes.c_float()
    assert L.debug_lib().ucd_selftest_umma(7, ctypes.byref(err)) == -1


def test_product_refuses_cpu_tensors(L):
    import ucd_b200
    x = torch.randn(1, 4, 8, 8)
    y = torch.zeros(1, 8, 8, dtype=torch.int64)
    with pytest.raises(RuntimeError, match="no CPU"):
        ucd_b200.UnbiasedCrossEntropy(old_cl=2)(x, y)
    with pytest.raises(RuntimeError, match="no CPU"):
        ucd_b200.UnbiasedKnowledgeDistillationLoss()(x, x[:, :2])
    with pytest.raises(RuntimeError, match="no CPU"):
        ucd_b200.interpolate_bilinear(x, (16, 16))
    f = torch.randn(1, 256, 4, 4)
    with pytest.raises(RuntimeError, match="no CPU"):
        ucd_b200.pre_contrastive_pixel(f, torch.zeros(1, 64, 64, dtype=torch.int64), l_po=torch.randn(1, 4, 4, 4), f_o=f)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No extension -> RuntimeError naming the build command; nothing falls back to PyTorch or the oracle."""
    from ucd_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libucd_b200.so"))
    with pytest.raises(RuntimeError, match="ucd_b200.build"):
        _lib.lib()
    import ucd_b200
    if torch.cuda.is_available():   # on a GPU box the modules themselves must raise, too
        with pytest.raises(RuntimeError, match="missing"):
            ucd_b200.interpolate_bilinear(torch.randn(1, 1, 4, 4, device="cuda"), (8, 8))


def test_payload_layout_and_views():
    """Host logic of the exchange payload: sections are aligned, disjoint and typed views alias the buffer."""
    from ucd_b200.losses import payload_layout, payload_views
    lay = payload_layout(49, 16)
    secs = sorted((lay[k], k) for k in ("counts", "range", "lab", "prob", "feat"))
    assert all(off % 256 == 0 for off, _ in secs) and secs[0] == (0, "counts")
    assert lay["nbytes"] >= 49 * (8 + 512 + 16 * 256 + 65536) and lay["nbytes"] % 256 == 0
    buf = torch.zeros(2, lay["nbytes"], dtype=torch.uint8)
    v = payload_views(buf, 49, 16)
    assert v["counts"].shape == (2, 4) and v["feat"].shape == (2, 49, 32, 128, 8) and v["prob"].shape == (2, 49, 2, 128, 8)
    assert v["lab"].shape == (2, 49, 128) and v["range"].shape == (2, 49, 2)
    v["counts"][1, 2] = 7
    v["feat"][1, 48, 31, 127, 7] = 1.0
    one = payload_views(buf[1], 49, 16)
    assert int(one["counts"][2]) == 7 and float(one["feat"][48, 31, 127, 7]) == 1.0
    assert int(buf[0].sum()) == 0


def test_sibling_and_opt_in_modules_exported():
    import inspect
    import ucd_b200 as U
    for name in ("KnowledgeDistillationLoss", "MaskKnowledgeDistillationLoss", "MaskCrossEntropy", "FusedUnbiasedLosses",
                 "PixelContrastiveDistillation"):
        assert inspect.isclass(getattr(U, name)), name
    assert list(inspect.signature(U.KnowledgeDistillationLoss.forward).parameters)[1:] == ["inputs", "targets", "mask"]
    assert list(inspect.signature(U.MaskCrossEntropy.forward).parameters)[1:] == ["inputs", "targets", "outputs_old"]
    assert list(inspect.signature(U.MaskCrossEntropy.__init__).parameters)[1:] == ["old_cl", "reduction", "ignore_index"]


def test_reference_signatures_preserved():
    """Constructor / forward signatures the trainer relies on (train.py:32,38,50,115-116,133)."""
    import inspect
    import ucd_b200 as U
    assert list(inspect.signature(U.UnbiasedCrossEntropy.__init__).parameters)[1:] == ["old_cl", "reduction", "ignore_index"]
    assert list(inspect.signature(U.UnbiasedKnowledgeDistillationLoss.__init__).parameters)[1:] == ["reduction", "alpha"]
    assert list(inspect.signature(U.PixelConLossV2.__init__).parameters)[1:3] == ["sample_method", "temperature"]
    assert list(inspect.signature(U.PixelConLossV2.forward).parameters)[1:] == [
        "anchor_features", "contrast_feature", "anchor_labels", "contrast_labels", "P"]
    assert list(inspect.signature(U.UnbiasedCrossEntropy.forward).parameters)[1:] == ["inputs", "targets"]
    assert list(inspect.signature(U.UnbiasedKnowledgeDistillationLoss.forward).parameters)[1:] == ["inputs", "targets", "mask"]
    assert list(inspect.signature(U.pre_contrastive_pixel).parameters)[:4] == ["f_n", "l_n", "l_po", "f_o"]
    assert U.pre_contractive_pixel is U.pre_contrastive_pixel
    m = U.PixelConLossV2(temperature=0.07)
    assert len(list(m.parameters())) == 0 and len(list(m.buffers())) == 0
    # the self-contrast siblings of utils/loss_new.py:263-400
    assert list(inspect.signature(U.PixelConLoss.__init__).parameters)[1:] == ["sample_method", "temperature"]
    assert list(inspect.signature(U.PixelConLoss.forward).parameters)[1:] == ["features", "labels"]
    assert list(inspect.signature(U.SupConLoss.__init__).parameters)[1:] == ["temperature", "contrast_mode", "base_temperature"]
    assert list(inspect.signature(U.SupConLoss.forward).parameters)[1:] == ["features", "labels", "mask"]
    assert U.PixelConLoss().temperature == 1 and U.SupConLoss().base_temperature == 0.07
    with pytest.raises(ValueError):   # same message path as loss_new.py:292-294: fewer than 3 dimensions
        U.SupConLoss()(torch.zeros(4, 8))


def test_no_product_import_of_oracle():
    """The product path must never import, include, load or execute anything under oracle/."""
    bad = re.compile(r"^\s*(from|import)\s+oracle|#include\s*[\"<][^\n]*oracle|(CDLL|open|exec|system|run)\([^\n]*oracle", re.M)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ucd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not bad.search(src), os.path.join(dirpath, f)
