"""Build the native library IN-TREE with nvcc for sm_100a (B200).

``python -m cnn_graph_b200.build`` or ``__graft_entry__.build()``.  The shared object is
``cnn_graph_b200/libcnn_graph_b200.so``; it is git-ignored but travels to the GPU box.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libcnn_graph_b200.so')

SOURCES = ['cg_graph.cu', 'cg_spmm.cu', 'cg_gemm.cu', 'cg_elementwise.cu', 'cg_filter.cu', 'cg_fused.cu', 'cg_dw_umma.cu', 'cg_dw_planes.cu', 'cg_dw_thin.cu', 'cg_clenshaw.cu', 'cg_contract_umma.cu', 'cg_gemm_umma.cu', 'cg_gemm_pipe.cu', 'cg_gemm_stream.cu', 'cg_thin.cu', 'cg_bmm.cu', 'cg_profile.cu', 'cg_umma_test.cu',
           'cg_host.cpp']

NVCC_FLAGS = (['-DCG_TRACE_BUILD'] if os.environ.get('CG_TRACE_BUILD') else []) + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fno-fast-math', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build_native(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link the C-ABI shared library."""
    nvcc = _nvcc()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(HERE, '..', 'include', 'cnn_graph_b200.h'))
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(objdir, os.path.splitext(src)[0] + '.o')
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', path, '-o', obj]
            if verbose:
                print(' '.join(cmd), flush=True)
            subprocess.check_call(cmd)
    if force or _stale(LIB, objs):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart_static',
                                                      '-lpthread', '-ldl', '-lrt']
        if verbose:
            print(' '.join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    print(build_native(force='--force' in sys.argv, verbose='-v' in sys.argv))
