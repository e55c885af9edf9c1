"""The C-ABI library loads without a GPU and exports every symbol include/council_b200.h declares."""
import os
import re

from council_gan_b200 import ops

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, 'include', 'council_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(cg_[a-z0-9_]+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(ops.LIB_PATH):
        from council_gan_b200 import build
        build.build()
    lib = ops.load_library()
    declared = header_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), 'libcouncil_b200.so does not export %s' % name
    assert set(declared) == set(ops.EXPORTS), (set(declared) ^ set(ops.EXPORTS))
    assert lib.cg_last_error() is not None


def test_product_has_no_cpu_path():
    """Constructing the CUDA op-set without a GPU must fail loudly (no silent fallback)."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        ops.CudaOps('cuda:0')
    from council_gan_b200 import Council_Trainer
    from common import config_for
    with pytest.raises(RuntimeError):
        Council_Trainer(config_for('glasses'), 'cuda:0')


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'council_gan_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert 'council_oracle' not in src and 'ops_torch' not in src, fn
