"""Who may use the oracle (test infrastructure): the product package never, bench.py only in its CPU legs."""
import ast
import glob
import os

from conftest import ROOT


def _oracle_imports(path):
    tree = ast.parse(open(path).read())
    found = []
    for node in ast.walk(tree):
        if isinstance(node, ast.ImportFrom) and node.module and node.module.split(".")[0] == "oracle":
            found += [(node.lineno, node.module, a.name) for a in node.names]
        elif isinstance(node, ast.Import):
            found += [(node.lineno, a.name, None) for a in node.names if a.name.split(".")[0] == "oracle"]
    return found


def test_product_package_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "dclip_b200", "**", "*.py"), recursive=True):
        assert not _oracle_imports(path), path
    assert not _oracle_imports(os.path.join(ROOT, "workloads.py"))


def test_bench_touches_the_oracle_only_for_the_cpu_arms():
    # the only oracle module bench.py may load is the loader of the staged reference / oracle port, and only inside the functions that
    # time the CPU side (`--impl reference`, every `cpu_baseline`)
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    assert not [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom)) and "oracle" in ast.dump(n)]
    for lineno, module, name in _oracle_imports(os.path.join(ROOT, "bench.py")):
        assert (module, name) == ("oracle", "reference_arm"), (lineno, module, name)
    src = open(os.path.join(ROOT, "bench.py")).read().splitlines()
    for lineno, _, _ in _oracle_imports(os.path.join(ROOT, "bench.py")):
        context = "\n".join(src[max(0, lineno - 6):lineno])
        assert "with_cpu" in context or "def cpu_baseline" in context or "rank != 0" in context, (lineno, context)
