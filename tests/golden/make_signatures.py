"""Freeze the signatures of the reference's hot-path functions (parsed with ast: audio_lib cannot be imported here).

    python tests/golden/make_signatures.py        # needs /root/reference (this container only)
"""
import ast
import json
import os

NAMES = ["calc_preemphasis", "calc_inv_preemphasis", "calc_PHN_target", "calc_MFCC_input", "griffin_lim_alg",
         "from_power_to_wav"]


def main():
    tree = ast.parse(open("/root/reference/audio_lib.py").read())
    out = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in NAMES:
            args = node.args.args
            defaults = [None] * (len(args) - len(node.args.defaults)) + list(node.args.defaults)
            out[node.name] = [[a.arg, "<required>" if d is None else repr(ast.literal_eval(d))] for a, d in zip(args, defaults)]
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_signatures.json")
    json.dump(out, open(path, "w"), indent=1, sort_keys=True)
    print(path, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
