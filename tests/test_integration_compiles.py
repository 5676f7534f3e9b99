"""The boundary as a C / C++ maintainer meets it: the Qt-side adapter of INTEGRATION.md is real
source (integration/fm-processor-b200.{h,cpp}) and must compile against the reference's OWN headers; a plain C
program (integration/c_caller.c) includes include/sdrjfm_b200.h, links libsdrjfm_b200.so and — on the GPU box —
runs the hot path.  No ctypes anywhere in this file's product-side calls."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
CALLER = os.path.join(ROOT, "integration", "c_caller")


def build_c_caller():
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-D_DEFAULT_SOURCE", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "integration", "c_caller.c"), "-L" + os.path.join(ROOT, "sdr-j-fm_b200"),
           "-lsdrjfm_b200", "-lm", "-Wl,-rpath," + os.path.join(ROOT, "sdr-j-fm_b200"), "-o", CALLER]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_adapter_compiles_against_the_reference_headers(pkg):
    """g++ -fsyntax-only of the fmProcessor replacement with the reference's device-handler.h, audiosink.h,
    rds-decoder.h, ringbuffer.h and fm-constants.h on the include path; Qt, libsndfile and PortAudio are
    replaced by the stand-ins of integration/qt_stub (none of them is installed here)."""
    if not os.path.isdir(os.path.join(REF, "includes")):
        pytest.skip("reference tree not present on this box")
    inc = [os.path.join(ROOT, "integration", "qt_stub"), os.path.join(ROOT, "include")]
    inc += [os.path.join(REF, d) for d in ("includes", "includes/various", "includes/fm", "includes/rds",
                                           "includes/output", "devices")]
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror"] + ["-I" + d for d in inc]
    cmd.append(os.path.join(ROOT, "integration", "fm-processor-b200.cpp"))
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_adapter_mirrors_every_public_method_of_the_reference(pkg):
    """every public method and signal the reference's fmProcessor declares (includes/fm/fm-processor.h:103-157,
    286-293) is declared by the adapter with the same name."""
    import re
    path = os.path.join(REF, "includes", "fm", "fm-processor.h")
    if not os.path.exists(path):
        pytest.skip("reference tree not present on this box")
    ref = open(path).read()
    pub = ref[ref.index("public:", ref.index("struct SMetaData")):ref.index("//	some private functions")]
    sig = ref[ref.rindex("signals:"):]
    names = set(re.findall(r"\b(\w+)\s*\(", pub)) | set(re.findall(r"\b(\w+)\s*\(", sig))
    names -= {"fmProcessor", "defined", "SMetaData"}
    mine = open(os.path.join(ROOT, "integration", "fm-processor-b200.h")).read()
    missing = sorted(n for n in names if not re.search(r"\b%s\s*\(" % re.escape(n), mine))
    assert not missing, missing


def test_c_program_links_the_library(pkg):
    """a C99 translation unit that includes the header and links -lsdrjfm_b200 builds without warnings"""
    pkg.lib()
    build_c_caller()
    assert os.path.exists(CALLER)


@pytest.mark.gpu
def test_c_program_runs_the_hot_path(pkg):
    """the C caller demodulates 0.25 s of a 1 kHz FM tone through sdrjfm_process and checks the audio itself"""
    pkg.lib()
    build_c_caller()
    r = subprocess.run([CALLER], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "1 kHz amplitude" in r.stdout
