#!/usr/bin/env python
"""Stamp a licence header on source files that lack one (counterpart of the reference's script/add-copyright.py)."""
import argparse
import pathlib

COMMENT = {".py": "# ", ".sh": "# ", ".cu": "// ", ".cuh": "// ", ".cpp": "// ", ".h": "// "}


def stamp(path: pathlib.Path, text: str, dry: bool) -> bool:
    prefix = COMMENT.get(path.suffix)
    if prefix is None:
        return False
    body = path.read_text()
    header = "".join(prefix + line + "\n" for line in text.splitlines())
    if text.splitlines()[0] in body[:400]:
        return False
    if not dry:
        if body.startswith("#!"):
            first, rest = body.split("\n", 1)
            path.write_text(first + "\n" + header + rest)
        else:
            path.write_text(header + body)
    return True


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("root", nargs="?", default=".")
    ap.add_argument("--text", default="Licensed under the Apache License, Version 2.0")
    ap.add_argument("--dry-run", action="store_true")
    a = ap.parse_args()
    n = 0
    for p in pathlib.Path(a.root).rglob("*"):
        if p.is_file() and ".git" not in p.parts and "_ref" not in p.parts and stamp(p, a.text, a.dry_run):
            n += 1
            print(("would stamp " if a.dry_run else "stamped ") + str(p))
    print(f"{n} file(s)")
