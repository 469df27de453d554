"""One-off pin of the drop-in boundary against the reference's own sources (needs /root/reference: build container only;
a script, not a collected test):   python tests/golden/pin_cli_and_config.py

  * every upper-case constant of the reference's src/config.py (importable without TensorFlow) equals ours;
  * every add_argument(...) of the reference's src/models/config_utils.py (parsed with ast, it imports TensorFlow) exists
    in ours with the same type / default / action expressions.
Last run (round 1): 26 / 26 constants equal; 19 / 19 flags equal; our extra flags: --adam-mode --reg-scale --plan-steps
--seed --device."""
import ast
import inspect
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"


def flags_of(source):
    out = {}
    for node in ast.walk(ast.parse(source)):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" and node.args:
            out[node.args[0].value] = {k.arg: ast.unparse(k.value) for k in node.keywords}
    return out


def main():
    sys.path.insert(0, ROOT)
    from glove_tensorflow_b200 import config as mine, config_utils
    scratch = tempfile.mkdtemp()
    for name in ("src", "configs"):
        os.symlink(os.path.join(REF, name), os.path.join(scratch, name))
    os.chdir(scratch)
    sys.path.insert(0, scratch)
    import src.config as ref
    names = [n for n in dir(ref) if n.isupper() and n != "CONFIG"]       # CONFIG is the ConfigParser section object itself
    bad = [(n, getattr(ref, n), getattr(mine, n, "<missing>")) for n in names if getattr(ref, n) != getattr(mine, n, "<missing>")]
    print("constants equal: %d / %d %s" % (len(names) - len(bad), len(names), bad or ""))
    theirs = flags_of(open(os.path.join(REF, "src", "models", "config_utils.py")).read())
    ours = flags_of(inspect.getsource(config_utils))
    diffs = [(f, k, theirs[f].get(k), ours.get(f, {}).get(k)) for f in theirs for k in ("type", "default", "action")
             if f not in ours or theirs[f].get(k) != ours[f].get(k)]
    print("flags equal: %d / %d %s; extra: %s" % (len(theirs) - len({d[0] for d in diffs}), len(theirs), diffs or "",
                                                  " ".join(f for f in ours if f not in theirs)))
    assert not bad and not diffs


if __name__ == "__main__":
    main()
