"""The prefix index's wire form: etcd key "XLLM:CACHE:" + 16 raw bytes, value =
CacheLocations::serialize_to_json().dump() (common/types.h:320-365, etcd_client.cpp:122-137,174-198).
host/index_wire.h is pinned against the REAL nlohmann::json the reference serialises with — the header ships in
this image under cudnn_frontend's thirdparty tree — by tests/cpp/index_wire_main.cc: byte-identical text for the
same name order, the reference's parser reads ours, ours reads the reference's (any order / pretty / \\u-escaped),
and both accept and reject the same malformed documents."""
import glob
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(__file__)
ROOT = os.path.dirname(HERE)


def _nlohmann_include():
    for base in sys.path + ["/usr/include", "/usr/local/include"]:
        for hit in glob.glob(os.path.join(base, "**", "nlohmann", "json.hpp"), recursive=True) if os.path.isdir(base) else []:
            return os.path.dirname(os.path.dirname(hit))
    return None


def test_wire_format_against_real_nlohmann(tmp_path):
    inc = _nlohmann_include()
    if inc is None:
        pytest.skip("no nlohmann/json.hpp in this image")
    exe = tmp_path / "index_wire_main"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", inc, os.path.join(HERE, "cpp", "index_wire_main.cc"),
                           "-o", str(exe)])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-2000:]


def test_listing_and_watch_semantics_against_a_stub(tmp_path):
    """host/index_snapshot.h without a GPU (the index entry points of the C-ABI are an in-memory stub): the PUTs of
    one etcd response are applied before its DELETEs, the last value of a key wins, unparsable pairs are skipped,
    and a snapshot reads back into an identical table."""
    exe = tmp_path / "index_snapshot_stub_main"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "xllm_service_b200", "host"),
                           os.path.join(HERE, "cpp", "index_snapshot_stub_main.cc"), "-o", str(exe)])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip().endswith("OK"), p.stdout[-1000:]
