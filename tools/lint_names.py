"""Tiny undefined-name check (no pyflakes offline): flags names that are read in a scope but never bound (at any point) in that
scope, an enclosing scope of the module, imported, or builtin.  Run before spending GPU minutes:  python tools/lint_names.py <files/dirs>"""
import ast
import builtins
import os
import sys


class V(ast.NodeVisitor):
    def __init__(self):
        self.scopes = [set()]
        self.uses = []  # (name, lineno, scope snapshot index)
        self.all_bound = set()

    def bind(self, name):
        self.scopes[-1].add(name)
        self.all_bound.add(name)

    def visit_FunctionDef(self, node):
        self.bind(node.name)
        for d in node.decorator_list:
            self.visit(d)
        for d in node.args.defaults + [k for k in node.args.kw_defaults if k is not None]:
            self.visit(d)
        self.scopes.append(set())
        a = node.args
        for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
            self.bind(x.arg)
        for s in node.body:
            self.visit(s)
        self.scopes.pop()

    visit_AsyncFunctionDef = visit_FunctionDef

    def visit_Lambda(self, node):
        self.scopes.append(set())
        a = node.args
        for x in a.posonlyargs + a.args + a.kwonlyargs + ([a.vararg] if a.vararg else []) + ([a.kwarg] if a.kwarg else []):
            self.bind(x.arg)
        self.visit(node.body)
        self.scopes.pop()

    def visit_ClassDef(self, node):
        self.bind(node.name)
        for b in node.bases + node.decorator_list:
            self.visit(b)
        self.scopes.append(set())
        for s in node.body:
            self.visit(s)
        self.scopes.pop()

    def visit_Import(self, node):
        for a in node.names:
            self.bind((a.asname or a.name).split(".")[0])

    def visit_ImportFrom(self, node):
        for a in node.names:
            self.bind(a.asname or a.name)

    def visit_Global(self, node):
        for n in node.names:
            self.bind(n)

    visit_Nonlocal = visit_Global

    def visit_ExceptHandler(self, node):
        if node.name:
            self.bind(node.name)
        self.generic_visit(node)

    def visit_Name(self, node):
        if isinstance(node.ctx, (ast.Store, ast.Del)):
            self.bind(node.id)
        else:
            self.uses.append((node.id, node.lineno, list(self.scopes)))

    def visit_comprehension(self, node):
        self.visit(node.target)
        self.visit(node.iter)
        for i in node.ifs:
            self.visit(i)

    def _comp(self, node):
        for g in node.generators:
            self.visit(g)
        for f in ("elt", "key", "value"):
            if hasattr(node, f):
                self.visit(getattr(node, f))

    visit_ListComp = visit_SetComp = visit_GeneratorExp = visit_DictComp = _comp

    def visit_MatchAs(self, node):
        if node.name:
            self.bind(node.name)
        self.generic_visit(node)


def check(path):
    src = open(path).read()
    try:
        tree = ast.parse(src, path)
    except SyntaxError as e:
        return [f"{path}:{e.lineno}: syntax error: {e.msg}"]
    v = V()
    v.visit(tree)
    known = set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    return [f"{path}:{ln}: undefined name '{n}'" for n, ln, chain in v.uses if n not in known and not any(n in sc for sc in chain)]


def main(args):
    files = []
    for a in args:
        if os.path.isdir(a):
            for d, _, fs in os.walk(a):
                files += [os.path.join(d, f) for f in fs if f.endswith(".py")]
        else:
            files.append(a)
    bad = []
    for f in sorted(files):
        bad += check(f)
    print("\n".join(bad) if bad else f"lint_names: {len(files)} files OK")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:] or ["afford-motion_b200", "tests", "tools", "oracle", "bench.py", "__graft_entry__.py"]))
