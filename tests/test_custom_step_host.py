"""kg_field2d_step_custom without a GPU: the generated CUDA source compiles for sm_100a (NVRTC runs on the host)."""
import ctypes as C
import ctypes.util

import pytest

from custom_models import BIRD_FINISH, BIRD_PAIR, CENTROID_FINISH, CENTROID_PAIR
from krabmaga_b200 import _abi as abi


def source(pair, finish, may_stop):
    need = abi.u64()
    abi.check(abi.lib().kg_jit_agent_source(pair.encode(), finish.encode(), int(may_stop), None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    abi.check(abi.lib().kg_jit_agent_source(pair.encode(), finish.encode(), int(may_stop), buf, need.value, C.byref(need)))
    return buf.value


def nvrtc():
    for name in ("libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            return C.CDLL(name)
        except OSError:
            pass
    pytest.skip("no NVRTC on this machine")


@pytest.mark.parametrize("pair,finish,may_stop", [(BIRD_PAIR, BIRD_FINISH, False), (CENTROID_PAIR, CENTROID_FINISH, True)])
def test_generated_source_compiles_for_sm_100a(pair, finish, may_stop):
    src = source(pair, finish, may_stop)
    assert b"kg_agent_step" in src and pair.strip().encode() in src
    rt = nvrtc()
    prog = C.c_void_p()
    assert rt.nvrtcCreateProgram(C.byref(prog), src, b"k.cu", 0, None, None) == 0
    opts = (C.c_char_p * 4)(b"--gpu-architecture=sm_100a", b"--fmad=false", b"--prec-div=true", b"--prec-sqrt=true")
    rc = rt.nvrtcCompileProgram(prog, 4, opts)
    n = C.c_size_t()
    rt.nvrtcGetProgramLogSize(prog, C.byref(n))
    log = C.create_string_buffer(max(n.value, 1))
    rt.nvrtcGetProgramLog(prog, log)
    assert rc == 0, log.value.decode(errors="replace")[-2000:]
    rt.nvrtcGetCUBINSize(prog, C.byref(n))
    assert n.value > 1000
    rt.nvrtcDestroyProgram(C.byref(prog))


def test_a_broken_snippet_is_reported_by_the_compiler():
    src = source("acc[0] +* 1;", "nx = sx;", False)
    rt = nvrtc()
    prog = C.c_void_p()
    assert rt.nvrtcCreateProgram(C.byref(prog), src, b"k.cu", 0, None, None) == 0
    opts = (C.c_char_p * 1)(b"--gpu-architecture=sm_100a")
    assert rt.nvrtcCompileProgram(prog, 1, opts) != 0
    rt.nvrtcDestroyProgram(C.byref(prog))
