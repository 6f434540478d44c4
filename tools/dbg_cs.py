import sys, os
sys.path.insert(0, os.getcwd())
import tests.test_gpu_explicit as T
import builtins
# monkeypatch: capture got/exp by re-running the body with a dump
src = open('tests/test_gpu_explicit.py').read()
body = src[src.index("def test_collapsed_stitched_full_text_golden_through_cuda_path"):]
body = body.replace("    assert len(got) == len(exp)\n    for a, b in zip(got, exp):\n        assert a == b\n", "    open('gpurun_out/cs_got.txt','w').write('\\n'.join(got))\n")
ns = dict(T.__dict__)
exec(body, ns)
ns['test_collapsed_stitched_full_text_golden_through_cuda_path']()
