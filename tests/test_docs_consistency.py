"""Every repository path the documents cite exists (README.md, profiles/README.md)."""
import glob
import os
import re

from conftest import ROOT

TOP = ("tests/", "profiles/", "scripts/", "oracle/", "openstaple_b200/", "include/")


def cited_paths(text):
    out = set()
    for tok in re.findall(r"`([^`\n]+)`", text):
        for part in re.split(r"[\s,;()]+", tok):
            part = part.split("::")[0].split(":")[0].rstrip(".")
            if part.startswith(TOP) and "<" not in part and "{" not in part and "…" not in part:
                out.add(part)
    return out


def test_cited_paths_exist():
    missing = []
    for doc in ("DESIGN.md", "INTEGRATION.md", "README.md", os.path.join("profiles", "README.md")):
        base = os.path.join(ROOT, "profiles") if doc.startswith("profiles") else ROOT
        text = open(os.path.join(ROOT, doc)).read()
        for p in cited_paths(text):
            if p.startswith("oracle/_ref"):          # built artefacts, git-ignored
                continue
            if not glob.glob(os.path.join(ROOT, p)) and not glob.glob(os.path.join(ROOT, p + "*")):
                missing.append((doc, p))
        if doc.startswith("profiles"):
            for tok in re.findall(r"`([^`\n]+)`", text):
                for part in re.split(r"[\s,;()]+", tok):
                    if re.match(r"^r0\d\w*_[\w.*{},]+$", part) and "{" not in part:
                        if not glob.glob(os.path.join(base, part)) and not glob.glob(os.path.join(base, part + "*")):
                            missing.append((doc, part))
    assert not missing, missing
