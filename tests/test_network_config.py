"""CPU: network config files — comments, "parent" inheritance as an RFC 7386 merge patch (Testbed::load_network_config +
merge_parent_network_config, src/testbed.cu:86-97, 280-309) — through the library's host-only hook ngp_load_network_config, which
ngp_testbed_reload_network_from_file shares.  Checked on the reference's own configs/ tree when /root/reference is present."""
import importlib
import json
import re
from pathlib import Path

import pytest

P = importlib.import_module("instant-ngp_b200")
CONFIGS = Path("/root/reference/configs")


def merge_patch(target, patch):
    """RFC 7386, written independently of the library"""
    if not isinstance(patch, dict):
        return patch
    if not isinstance(target, dict):
        target = {}
    out = dict(target)
    for k, v in patch.items():
        if v is None:
            out.pop(k, None)
        else:
            out[k] = merge_patch(out.get(k), v)
    return out


def load_independently(path):
    text = re.sub(r"//[^\n]*|/\*.*?\*/", "", Path(path).read_text(), flags=re.S)
    child = json.loads(text)
    if "parent" not in child:
        return child
    return merge_patch(load_independently(Path(path).parent / child["parent"]), child)


def test_parent_chain_null_deletion_and_comments(tmp_path):
    (tmp_path / "base.json").write_text('''{
        // the root
        "loss": {"otype": "Huber"},
        "encoding": {"otype": "HashGrid", "n_levels": 16, "log2_hashmap_size": 19, "extra": 1},
        "list": [1, 2, 3]
    }''')
    (tmp_path / "mid.json").write_text('{"parent": "base.json", "encoding": {"log2_hashmap_size": 15, "extra": null}, /* replaced, not merged */ "list": [9]}')
    sub = tmp_path / "sub"
    sub.mkdir()
    (sub / "leaf.json").write_text('{"parent": "../mid.json", "loss": {"otype": "L2"}, "new": {"a": {"b": 1}}}')
    got = P.load_network_config(sub / "leaf.json")
    assert got == {"loss": {"otype": "L2"}, "encoding": {"otype": "HashGrid", "n_levels": 16, "log2_hashmap_size": 15}, "list": [9], "new": {"a": {"b": 1}},
                   "parent": "../mid.json"}                      # the child's own "parent" key stays, as nlohmann's merge_patch leaves it
    assert got == load_independently(sub / "leaf.json")
    assert P.load_network_config(tmp_path / "base.json")["list"] == [1, 2, 3]


def test_errors(tmp_path):
    (tmp_path / "orphan.json").write_text('{"parent": "missing.json"}')
    with pytest.raises(P.NgpError, match="does not exist"):
        P.load_network_config(tmp_path / "orphan.json")
    (tmp_path / "loop.json").write_text('{"parent": "loop.json"}')
    with pytest.raises(P.NgpError, match="too deep"):
        P.load_network_config(tmp_path / "loop.json")
    with pytest.raises(P.NgpError, match="does not exist"):
        P.load_network_config(tmp_path / "nothing.json")


@pytest.mark.skipif(not CONFIGS.exists(), reason="reference configs not present (GPU box)")
def test_every_reference_config_resolves_like_merge_patch():
    files = sorted(CONFIGS.glob("*/*.json"))
    assert len(files) >= 20
    n_children = 0
    for f in files:
        got = P.load_network_config(f)
        assert got == load_independently(f), f
        n_children += "parent" in got
    assert n_children >= 8
    big = P.load_network_config(CONFIGS / "nerf" / "big.json")
    base = P.load_network_config(CONFIGS / "nerf" / "base.json")
    rest = lambda e: {k: v for k, v in e.items() if k != "log2_hashmap_size"}   # noqa: E731
    assert big["encoding"]["log2_hashmap_size"] == 21 and rest(big["encoding"]) == rest(base["encoding"]) and big["network"] == base["network"]
    two = P.load_network_config(CONFIGS / "nerf" / "base_2layer.json")    # parent linear.json, whose parent is base.json
    assert two["rgb_network"]["n_hidden_layers"] == 2 and two["optimizer"] == P.load_network_config(CONFIGS / "nerf" / "linear.json")["optimizer"]
