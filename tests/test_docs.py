"""The evidence the documents cite exists: every profiles/ file named in profiles/README.md,
DESIGN.md and README.md is in the tree (wildcards must match something)."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cited(path, pattern):
    text = open(os.path.join(ROOT, path)).read()
    return sorted(set(re.findall(pattern, text)))


def exists(relative):
    relative = relative.replace("NN", "*")
    return bool(glob.glob(os.path.join(ROOT, relative)))


def test_profiles_index_names_existing_files():
    names = cited("profiles/README.md", r"`((?:r0\d|traffic)[\w*.\-]+)`")
    assert len(names) > 20
    missing = [name for name in names if not exists(os.path.join("profiles", name))]
    assert not missing, missing


def test_documents_cite_existing_evidence_and_sources():
    for document in ("DESIGN.md", "README.md", "INTEGRATION.md"):
        paths = cited(document, r"`((?:profiles|scripts|tests|oracle|include|embiggen_b200)/[\w*./\-]+)`")
        # named on purpose although absent: the reference's own fixture, and the oracle/_ref that
        # cannot exist here (the reference's engine is an un-vendored wheel)
        allowed = {"tests/data/small_ppi.tsv", "oracle/_ref"}
        missing = [p for p in paths if p.rstrip(".") not in allowed and not exists(p.rstrip("."))]
        assert not missing, (document, missing)
