"""CPU-side checks of the drop-in boundary: the shared library loads and exports exactly what include/isaac_ext.h
declares; struct layouts agree; compute without a device fails loudly (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "isaac_ext.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isaac_ext_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from isaac_aligner_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    declared = header_functions()
    assert declared, "no functions parsed from the header"
    for name in declared:
        assert hasattr(lib, name), "libisaac_ext.so does not export " + name
    assert sorted(capi.EXPORTS) == declared


def test_struct_layouts_match_header():
    from isaac_aligner_b200.types import CANDIDATE_DTYPE, FRAGMENT_DTYPE, Config, Reads
    assert FRAGMENT_DTYPE.itemsize == 64 and CANDIDATE_DTYPE.itemsize == 16
    assert ctypes.sizeof(Config) == 13 * 4
    assert ctypes.sizeof(Reads) == 6 * 4 + 2 * 8
    assert FRAGMENT_DTYPE.fields["matchCount"][1] == 62 and FRAGMENT_DTYPE.fields["observedLength"][1] == 32


def test_no_cpu_fallback():
    """Without a CUDA device the library must refuse to compute instead of silently using the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from isaac_aligner_b200 import capi
    from isaac_aligner_b200.types import Config
    with pytest.raises(capi.ExtError) as e:
        capi.Context(Config.default())
    assert e.value.code == 2   # ISAAC_EXT_E_NO_DEVICE


def test_product_does_not_touch_the_oracle():
    """Nothing under isaac_aligner_b200/ may include, link or load anything under oracle/."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "isaac_aligner_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"oracle[_/]|libisaac_oracle|libisaac_ref|oracle_api", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_python_mirrors_have_the_c_struct_sizes():
    """every ctypes.Structure / numpy dtype the harness passes over the ABI against sizeof() of the header's struct (gcc)"""
    import ctypes
    import subprocess
    from isaac_aligner_b200 import batch, bins, synth, tile, types
    pairs = [("isaac_ext_tile_t", ctypes.sizeof(tile.TileC)), ("isaac_ext_tile_result_t", ctypes.sizeof(tile.TileResultC)),
             ("isaac_ext_config_t", ctypes.sizeof(types.Config)), ("isaac_ext_reads_t", ctypes.sizeof(types.Reads)),
             ("isaac_ext_adapter_t", ctypes.sizeof(types.Adapter)), ("isaac_ext_fragment_t", types.FRAGMENT_DTYPE.itemsize),
             ("isaac_ext_candidate_t", types.CANDIDATE_DTYPE.itemsize), ("isaac_ext_alignment_t", types.ALIGNMENT_DTYPE.itemsize), ("isaac_ext_match_t", synth.MATCH_DTYPE.itemsize),
             ("isaac_ext_seed_t", synth.SEED_DTYPE.itemsize), ("isaac_ext_build_batch_t", ctypes.sizeof(batch.BuildBatch)),
             ("isaac_ext_build_result_t", ctypes.sizeof(batch.BuildResult)), ("isaac_ext_tls_t", ctypes.sizeof(batch.Tls)),
             ("isaac_ext_rescue_request_t", batch.RESCUE_REQUEST_DTYPE.itemsize), ("isaac_ext_rescue_result_t", ctypes.sizeof(batch.RescueResult)),
             ("isaac_ext_template_options_t", ctypes.sizeof(batch.TemplateOptions)), ("isaac_ext_template_t", batch.TEMPLATE_DTYPE.itemsize),
             ("isaac_ext_template_result_t", ctypes.sizeof(batch.TemplateResult)), ("isaac_ext_pack_options_t", ctypes.sizeof(batch.PackOptionsC)),
             ("isaac_ext_pack_result_t", ctypes.sizeof(batch.PackResultC)), ("isaac_ext_bin_index_t", bins.BIN_INDEX_DTYPE.itemsize),
             ("isaac_ext_gap_t", bins.GAP_DTYPE.itemsize), ("isaac_ext_realign_options_t", ctypes.sizeof(bins.RealignOptionsC)),
             ("isaac_ext_realign_result_t", ctypes.sizeof(bins.RealignResultC)), ("isaac_ext_realign_job_t", ctypes.sizeof(bins.RealignJobC))]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "build", "abi_sizes.c")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    with open(src, "w") as f:
        f.write('#include <stdio.h>\n#include "isaac_ext.h"\nint main(void) {\n')
        for name, _ in pairs:
            f.write('    printf("%s %%zu\\n", sizeof(%s));\n' % (name, name))
        f.write("    return 0;\n}\n")
    exe = os.path.join(root, "build", "abi_sizes")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), src, "-o", exe])
    out = dict(line.split() for line in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, size in pairs:
        assert int(out[name]) == size, (name, out[name], size)


def test_bin_record_layouts_agree():
    """io::FragmentHeader has four statements here: the host mirror (host/isaac_b200.hh), the realigner's byte offsets
    (csrc/realign_device.cuh), the packer's struct (csrc/pack_fragments.cuh) and the harness dtype (bins.HEADER_DTYPE)"""
    import subprocess
    from isaac_aligner_b200 import bins
    src = os.path.join(ROOT, "build", "bin_layout.cu")
    os.makedirs(os.path.dirname(src), exist_ok=True)
    fields = ["bamTlen", "observedLength", "fStrandPosition", "lowClipped", "highClipped", "alignmentScore", "templateAlignmentScore",
              "mateFStrandPosition", "readLength", "cigarLength", "gapCount", "editDistance", "flags", "tile", "barcode", "barcodeSequence",
              "clusterId", "clusterX", "clusterY", "duplicateClusterRank", "mateAnchor", "mateStorageBin"]
    with open(src, "w") as f:
        f.write('#include <cstdio>\n#include <cstddef>\n#include "../isaac_aligner_b200/host/isaac_b200.hh"\n'
                '#include "../isaac_aligner_b200/csrc/realign_device.cuh"\n#include "../isaac_aligner_b200/csrc/pack_fragments.cuh"\n'
                'int main() {\n  using isaac_b200::io::FragmentHeader; using isaac_b200::PackedFragmentHeader;\n')
        for name in fields:
            f.write('  std::printf("%s %%zu %%zu\\n", offsetof(FragmentHeader, %s_), offsetof(PackedFragmentHeader, %s));\n' % (name, name, name))
        f.write('  std::printf("sizeof %zu %zu\\n", sizeof(FragmentHeader), sizeof(PackedFragmentHeader));\n')
        for name in ("BAM_TLEN", "OBSERVED_LENGTH", "F_STRAND_POSITION", "LOW_CLIPPED", "HIGH_CLIPPED", "ALIGNMENT_SCORE", "TEMPLATE_ALIGNMENT_SCORE",
                     "MATE_F_STRAND_POSITION", "READ_LENGTH", "CIGAR_LENGTH", "GAP_COUNT", "EDIT_DISTANCE", "FLAGS", "BARCODE", "CLUSTER_ID"):
            f.write('  std::printf("BIN_%s %%u\\n", unsigned(isaac_b200::BIN_%s));\n' % (name, name))
        f.write("  return 0;\n}\n")
    exe = os.path.join(ROOT, "build", "bin_layout")
    subprocess.check_call(["nvcc", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "include"), src, "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split("\n")
    rows = dict((line.split()[0], [int(x) for x in line.split()[1:]]) for line in out if line.strip())
    for name in fields:
        want = bins.HEADER_DTYPE.fields[name][1]
        assert rows[name] == [want, want], (name, rows[name], want)
    assert rows["sizeof"] == [112, 112] and bins.HEADER_DTYPE.itemsize == 112
    camel = {"BAM_TLEN": "bamTlen", "OBSERVED_LENGTH": "observedLength", "F_STRAND_POSITION": "fStrandPosition", "LOW_CLIPPED": "lowClipped",
             "HIGH_CLIPPED": "highClipped", "ALIGNMENT_SCORE": "alignmentScore", "TEMPLATE_ALIGNMENT_SCORE": "templateAlignmentScore",
             "MATE_F_STRAND_POSITION": "mateFStrandPosition", "READ_LENGTH": "readLength", "CIGAR_LENGTH": "cigarLength", "GAP_COUNT": "gapCount",
             "EDIT_DISTANCE": "editDistance", "FLAGS": "flags", "BARCODE": "barcode", "CLUSTER_ID": "clusterId"}
    for key, name in camel.items():
        assert rows["BIN_" + key] == [bins.HEADER_DTYPE.fields[name][1]], key
