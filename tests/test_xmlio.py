"""toXML / fromXML in the reference's XML dialect (AbstractHMM.java:394-465, XMLParser.java:233-271)."""
import numpy as np
import pytest

from helpers import order0_hmm, silent_hmm, sparse_hmm
from jstacs_b200 import HigherOrderHMM
from jstacs_b200.workloads import banded_hmm, ergodic_hmm
from jstacs_b200.xmlio import append_object, java_double, parse, to_object


def test_java_double_to_string():
    cases = {0.0: "0.0", 1.0: "1.0", -1.3862943611198906: "-1.3862943611198906", 100.0: "100.0", 1234567.0: "1234567.0",
             1.0e7: "1.0E7", 12345678.9: "1.23456789E7", 0.001: "0.001", 0.00099: "9.9E-4", 1e-5: "1.0E-5", 2.5e-17: "2.5E-17",
             float("inf"): "Infinity", float("-inf"): "-Infinity", 0.25: "0.25", 123.456: "123.456", 1e22: "1.0E22"}
    for v, s in cases.items():
        assert java_double(v) == s, (v, java_double(v), s)
        if np.isfinite(v):
            assert float(java_double(v)) == v
    assert java_double(float("nan")) == "NaN" and java_double(-0.0) == "-0.0"
    rng = np.random.default_rng(0)
    for v in np.concatenate([rng.standard_normal(200) * 10.0 ** rng.integers(-30, 30, 200), -np.log(rng.random(50))]):
        assert float(java_double(v)) == v  # shortest round-trip digits survive the reformatting


def test_xmlparser_syntax():
    assert append_object("x", None) == "<x>null</x>"
    assert append_object("n", 3) == "<n><className>java.lang.Integer</className>3</n>"
    assert append_object("b", True) == "<b><className>java.lang.Boolean</className>true</b>"
    assert append_object("a", [1, 2], "[I") == '<a><className>[I</className><length>2</length><pos val="0">1</pos><pos val="1">2</pos></a>'
    m = append_object("m", [[0.5], [1.0, 2.0]], "[[D")
    assert m == ('<m><className>[[D</className><length>2</length><pos val="0"><className>[D</className><length>1</length><pos val="0">0.5</pos></pos>'
                 '<pos val="1"><className>[D</className><length>2</length><pos val="0">1.0</pos><pos val="1">2.0</pos></pos></m>')
    assert to_object(parse(m).child("m")) == [[0.5], [1.0, 2.0]]
    s = append_object("s", ["a<b", "c&d"], "[Ljava.lang.String;")
    assert "a&lt;b" in s and to_object(parse(s).child("s")) == ["a<b", "c&d"]
    assert to_object(parse(append_object("e", [], "[[I")).child("e")) == []


MODELS = {"ergodic_s4_k2": lambda: ergodic_hmm(4, 2, ess=3.0), "sparse_unsorted": sparse_hmm, "silent_profile": silent_hmm,
          "order0": lambda: order0_hmm(ess=1.0), "banded_s64": lambda: banded_hmm(64, 3, 1, offsets=[1, -1, 0])}


@pytest.mark.parametrize("name", list(MODELS))
def test_model_round_trip_is_exact(name):
    hmm = MODELS[name]()
    # -inf parameters (zero counts without a prior) survive
    if name == "sparse_unsorted":
        p = hmm.emission[0].params.copy()
        p[0, 0] = -np.inf
        hmm.emission[0].setLogProbabilities(p)
    xml = hmm.toXML()
    for tag in ("trainingParameter", "transition", "name", "filter", "emissionIdx", "strand", "emission", "statesGroups", "skipInit", "type"):
        assert f"<{tag}>" in xml  # AbstractHMM.java:396-411, HigherOrderHMM.java:229-230
    assert xml.startswith("<HigherOrderHMM>\n") and xml.endswith("</HigherOrderHMM>\n")
    assert "TRANSITION_ELEMENT" in xml and ("ConditionalDiscreteEmission" in xml)
    back = HigherOrderHMM.fromXML(xml)
    assert type(back) is type(hmm) and type(back.transition) is type(hmm.transition)
    a, b = hmm.ir(), back.ir()
    for f in a.FIELDS_I32 + a.FIELDS_U8:
        assert np.array_equal(getattr(a, f), getattr(b, f)), f
    for f in ("trans_param", "trans_hyper", "emis_param", "emis_hyper"):
        assert np.array_equal(getattr(a, f), getattr(b, f)), f  # doubles survive bit for bit
    # logNorm is recomputed on load, like the reference's fromXML -> precompute()
    np.testing.assert_allclose(b.trans_lognorm, a.trans_lognorm, atol=1e-15)
    assert back.name == hmm.name and back.trainingParameter.getNumberOfStarts() == hmm.trainingParameter.getNumberOfStarts()
    assert type(back.trainingParameter.getTerminationCondition()) is type(hmm.trainingParameter.getTerminationCondition())
    assert back.toXML() == xml  # idempotent


@pytest.mark.gpu
def test_trained_parameters_survive_the_xml_round_trip():
    import __graft_entry__ as g

    g.build()
    from jstacs_b200.workloads import planted_dna

    hmm = ergodic_hmm(3, 1, n_iter=4, ess=2.0)
    data = planted_dna(50, 100)
    hmm.train(data)
    back = HigherOrderHMM.fromXML(hmm.toXML())
    np.testing.assert_allclose(back.getLogScoreFor(data), hmm.getLogScoreFor(data), rtol=1e-13)
    for (p1, s1), (p2, s2) in zip(hmm.getViterbiPathsFor(data), back.getViterbiPathsFor(data)):
        assert np.array_equal(p1, p2)
