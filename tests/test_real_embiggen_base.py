"""The `HAVE_EMBIGGEN = True` branch: the B200 embedders under the reference's OWN base classes.

`embiggen_b200/embedding_api.py` subclasses the real `AbstractEmbeddingModel` whenever `embiggen`
imports, and a restatement otherwise; only the restatement ever ran in round 1.  Here the real
package is imported from /root/reference (third-party packages that are not installed are stubbed,
see tests/real_embiggen_probe.py) and the reference's own code checks our classes: the
constructor-time "no useless method" cross-checks that read the source of every capability
method (abstract_model.py:58-131), `parameters()` round trips, `into_smoke_test`, the registry
(`get_model_from_library(..., library_name="B200")`), `fit_transform`'s graph validation against
the duck-typed CSR graph, the real `EmbeddingResult`, and the real `embed_graph` resolving library
"B200" and re-raising the engine's "no CUDA device" error as ValueError.  Needs the reference
tree, so it runs in the build container and skips on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isdir("/root/reference/embiggen"), reason="the reference tree is not on this machine")
def test_embedders_satisfy_the_real_embiggen_base_classes():
    done = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "real_embiggen_probe.py")],
                          capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert done.returncode == 0, done.stderr[-3000:]
    line = [l for l in done.stdout.splitlines() if l.startswith("REAL_EMBIGGEN_OK ")]
    assert line, done.stdout[-2000:]
    report = json.loads(line[0][len("REAL_EMBIGGEN_OK "):])
    assert report["have_embiggen"] is True
    assert report["registered"] == ["DeepWalk CBOW", "DeepWalk SkipGram", "Node2Vec CBOW", "Node2Vec SkipGram"]
    assert "Walklets SkipGram" in report["models"] and "Node2Vec GloVe" in report["models"]
    assert "ensmallen" in report["stubs"]  # the engine behind the reference classes is what is absent
    # the reference's constructor cross-checks and capability defaults, case by case: the outcomes
    # the restatement is held to in tests/test_embedder_api.py are the reference's own
    import capability_cases
    assert report["capability_cases"] == capability_cases.expected_outcomes()
    import validation_cases
    assert report["validation_cases"] == validation_cases.expected_outcomes()
    import embed_graph_cases
    assert report["embed_graph_cases"] == embed_graph_cases.EXPECTED
    import embedding_result_cases
    assert report["embedding_result_cases"] == embedding_result_cases.EXPECTED
    # the adapter classes themselves: signature, defaults, parameters(), smoke conversion, names and
    # capability answers of the reference's eight classes against ours (computed here, under the
    # restated base classes), equal once the keyword-only B200 extras are left out
    import adapter_cases
    from embiggen_b200 import embedders
    ours = adapter_cases.describe({
        "Node2Vec SkipGram": embedders.Node2VecSkipGramB200, "Node2Vec CBOW": embedders.Node2VecCBOWB200,
        "DeepWalk SkipGram": embedders.DeepWalkSkipGramB200, "DeepWalk CBOW": embedders.DeepWalkCBOWB200,
        "Walklets SkipGram": embedders.WalkletsSkipGramB200, "Walklets CBOW": embedders.WalkletsCBOWB200,
        "Node2Vec GloVe": embedders.Node2VecGloVeB200, "DeepWalk GloVe": embedders.DeepWalkGloVeB200,
    }, drop=set(embedders._B200_DEFAULTS))
    theirs = report["adapter_description"]
    assert sorted(theirs) == sorted(ours)
    for name in theirs:
        for section in theirs[name]:
            assert json.loads(json.dumps(ours[name][section])) == theirs[name][section], (name, section)
    # the registry table: the reference's get_models_dataframe over our classes against the restated one
    from embiggen_b200.embedding_api import get_models_dataframe
    frame = get_models_dataframe()
    frame = frame[frame.library_name == "B200"].sort_values("model_name")
    assert json.loads(frame.to_json(orient="records")) == report["registry_rows"] and len(frame) == 4
    # ... and the edge-prediction perceptron of the step after the path (perceptron.py:15-300)
    from embiggen_b200.edge_prediction import PerceptronEdgePredictionB200
    ours = adapter_cases.describe_perceptron(PerceptronEdgePredictionB200, drop={"device"})
    assert json.loads(json.dumps(ours)) == report["perceptron_description"]
