"""The TensorFlow custom-op shim (deepsphere-cosmo-tf2_b200/tf_shim/) is source only in this image (no TensorFlow,
SURVEY F4).  What can be checked without TF: every C-ABI call it makes exists in include/deepsphere_b200.h with the
same number of arguments.  The load test runs only where TensorFlow is installed."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "deepsphere-cosmo-tf2_b200", "tf_shim")
HEADER = open(os.path.join(ROOT, "include", "deepsphere_b200.h")).read()


def _split_args(s):
    depth, cur, out = 0, "", []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def _calls(src, name):
    res = []
    for m in re.finditer(r"\b" + name + r"\s*\(", src):
        depth, i = 1, m.end()
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        res.append(src[m.end(): i - 1])
    return res


def test_every_c_abi_call_of_the_shim_matches_the_header():
    src = open(os.path.join(SHIM, "deepsphere_tf_ops.cc")).read()
    src = re.sub(r"//[^\n]*", "", src)
    header = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    used = sorted(set(re.findall(r"\b(ds_[a-z0-9_]+)\s*\(", src)))
    assert {"ds_graph_conv_forward", "ds_graph_conv_backward", "ds_pool_forward", "ds_bn_stats",
            "ds_bn_bias_act_forward"} <= set(used)
    for name in used:
        decl = _calls(header, name)
        assert len(decl) == 1, f"{name} is not declared exactly once in the header"
        n_decl = len(_split_args(decl[0])) if decl[0].strip() not in ("", "void") else 0
        for call in _calls(src, name):
            assert len(_split_args(call)) == n_decl, (name, call)


def test_shim_loads_under_tensorflow():
    pytest.importorskip("tensorflow")
    if not os.path.exists(os.path.join(SHIM, "deepsphere_tf_ops.so")):
        pytest.skip("deepsphere_tf_ops.so not built (see the build line at the top of deepsphere_tf_ops.cc)")
    import importlib.util

    spec = importlib.util.spec_from_file_location("deepsphere_tf", os.path.join(SHIM, "deepsphere_tf.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert hasattr(mod, "graph_conv")
