"""Malformed `ser/` directories must come back as VSGPU_EIO / VSGPU_ESHAPE errors, never as a crash
or a silently wrong index (loader + flattener run identically under the CUDA library and the
test-only host build)."""
import os
import shutil

import pytest

import vs_testlib as T


def open_hostsim(prefix):
    from variantstore_b200 import VariantStoreIndex, load_library
    return VariantStoreIndex(prefix, lib=load_library(T.HOSTSIM_SO, subset=True))


@pytest.fixture()
def ser_copy(tmp_path):
    dst = str(tmp_path / "ser")
    shutil.copytree(os.path.join(T.GOLDEN, "x_ser"), dst)
    return dst


@pytest.mark.parametrize("victim", ["index.sdsl", "ref_node_id.sdsl", "adj_list.cqf", "aux_vertex_list.sdsl", "seq_buffer.sdsl",
                                    "sample_vector.sdsl", "vertex_list_0.proto", "sampleid_map.lst"])
def test_missing_file(ser_copy, victim):
    from variantstore_b200 import VsgpuError
    os.remove(os.path.join(ser_copy, victim))
    with pytest.raises(VsgpuError) as ei:
        open_hostsim(ser_copy)
    assert ei.value.code == -2


@pytest.mark.parametrize("victim", ["index.sdsl", "adj_list.cqf", "vertex_list_0.proto", "seq_buffer.sdsl", "sample_vector.sdsl"])
def test_truncated_file(ser_copy, victim):
    from variantstore_b200 import VsgpuError
    p = os.path.join(ser_copy, victim)
    data = open(p, "rb").read()
    open(p, "wb").write(data[: len(data) // 2])
    with pytest.raises(VsgpuError) as ei:
        open_hostsim(ser_copy)
    assert ei.value.code in (-2, -3)


def test_bad_cqf_magic(ser_copy):
    from variantstore_b200 import VsgpuError
    p = os.path.join(ser_copy, "adj_list.cqf")
    data = bytearray(open(p, "rb").read())
    data[0] ^= 0xFF
    open(p, "wb").write(bytes(data))
    with pytest.raises(VsgpuError) as ei:
        open_hostsim(ser_copy)
    assert ei.value.code == -2 and "magic" in str(ei.value)


def test_index_of_another_graph_is_rejected(tmp_path, ser_copy):
    """index.sdsl / ref_node_id.sdsl that do not describe this graph's ref path -> VSGPU_ESHAPE."""
    from variantstore_b200 import VsgpuError
    other = os.path.join(T.GOLDEN, "xsmall_ser")
    for f in ("index.sdsl", "ref_node_id.sdsl"):
        shutil.copy(os.path.join(other, f), os.path.join(ser_copy, f))
    with pytest.raises(VsgpuError) as ei:
        open_hostsim(ser_copy)
    assert ei.value.code == -3


def test_sample_count_mismatch(ser_copy):
    from variantstore_b200 import VsgpuError
    p = os.path.join(ser_copy, "sampleid_map.lst")
    lines = open(p).read().split("\n")
    open(p, "w").write("\n".join(lines[:2] + lines[3:]))       # drop one "<name> <id>" row
    with pytest.raises(VsgpuError):
        open_hostsim(ser_copy)


def test_sample_coordinate_operators_need_the_vertex_blocks_again(ser_copy):
    """t3 / t5 read the per-carrier sample positions by a second pass over vertex_list_<k>.proto on
    their first call: blocks that vanished or changed since vsgpu_open are an error, not a wrong answer;
    t2 (which needs nothing new from disk) keeps working."""
    import numpy as np
    from variantstore_b200 import VsgpuError
    e = open_hostsim(ser_copy)
    assert e.query_sample_from_ref(10, 20, "1") == "TTTGAAAATT"
    victim = os.path.join(ser_copy, "vertex_list_0.proto")
    os.rename(victim, victim + ".away")
    for call in (lambda: e.batch_sample_seq_in_sample([10], [20], [1]), lambda: e.batch_sample_var_in_sample([10], [20], [1])):
        with pytest.raises(VsgpuError) as ei:
            call()
        assert ei.value.code == -3
    assert e.query_sample_from_ref(10, 20, "1") == "TTTGAAAATT"
    os.rename(victim + ".away", victim)
    off, text, st, _ = e.batch_sample_seq_in_sample([1], [1001], [1])      # the tables are built now
    assert st[0] == 0 and len(text) == 1000
    off5, hits5, st5, _ = e.batch_sample_var_in_sample([1], [1001], [1])
    assert st5[0] == 0 and len(hits5) == 74
    with pytest.raises(VsgpuError):
        e.batch_sample_var_in_sample([1], [1001], [2])                     # sample id out of range (one sample)
    e.close()


def test_short_but_well_formed_seq_buffer(ser_copy):
    """seq_buffer.sdsl cut to a twentieth of its symbols with a coherent header (int_vector<0>: u64 size in bits, u8
    width, words): every vertex slice past the new end must be refused at open — the materialiser and the
    render / copy kernels read those slices unchecked."""
    import struct
    from variantstore_b200 import VsgpuError
    p = os.path.join(ser_copy, "seq_buffer.sdsl")
    data = open(p, "rb").read()
    bits, width = struct.unpack_from("<QB", data)
    keep = (bits // width) // 20
    nwords = (keep * width + 63) // 64
    open(p, "wb").write(struct.pack("<QB", keep * width, width) + data[9:9 + 8 * nwords])
    with pytest.raises(VsgpuError) as ei:
        open_hostsim(ser_copy)
    assert ei.value.code == -2 and "beyond seq_buffer" in str(ei.value)


@pytest.mark.parametrize("field,value", [(32, 1 << 40), (40, 1 << 40), (72, 0), (64, 0), (48, 13), (96, 1 << 50)])
def test_cqf_header_fields_are_validated(ser_copy, field, value):
    """nslots / xnslots / bits_per_slot / key_remainder_bits / key_bits / nblocks of the qfmetadata header (gqf_int.h:84-101)
    drive every slot read of the scan; values that would index past the mapped blocks are refused."""
    import struct
    from variantstore_b200 import VsgpuError
    p = os.path.join(ser_copy, "adj_list.cqf")
    data = bytearray(open(p, "rb").read())
    struct.pack_into("<Q", data, field, value)
    open(p, "wb").write(bytes(data))
    with pytest.raises(VsgpuError) as ei:
        open_hostsim(ser_copy)
    assert ei.value.code == -2
