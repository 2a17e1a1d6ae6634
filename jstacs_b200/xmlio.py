"""Persistence in the reference's XML dialect: AbstractHMM.toXML / fromXML (HMM/AbstractHMM.java:394-465),
written and read with the syntax of de/jstacs/io/XMLParser.java:233-271 (className / length / pos val="i" / null).

Scope.  Everything that belongs to the HMM hot path is written tag for tag as Jstacs writes it and read back from either
side's files: <transition> (BasicHigherOrderTransition.java:130-137: transitions, isSilent, transIndex; element
:881-891: context, states, hyperparameters, parameters, weight, norm; HigherOrderTransition.java:126-134: offset),
<name>, <filter>, <emissionIdx>, <strand>, <emission> (AbstractConditionalDiscreteEmission.java:484-524: params, norm,
offset, hyperParams, initHyperParams, statistic, ess, shape, linear, hasParameters, order), <statesGroups>, <skipInit>,
<type>, <ess> (HigherOrderHMM.java:228-231, DifferentiableHigherOrderHMM.java:211-214).  The <trainingParameter> element
belongs to the reference's generic parameter framework (de/jstacs/parameters, outside the hot path): it is written as a
compact parameterSet holding the values only (starts, termination condition, threads), and when a Jstacs-written file is
read its values are recovered by tag search, defaults otherwise.  The image has no JVM, so the byte-level form could not
be diffed against a file produced by Jstacs; the round trip through this module is tested instead.
"""
from __future__ import annotations

import math
import re

import numpy as np

HMM_PKG = "de.jstacs.sequenceScores.statisticalModels.trainable.hmm"
SIMPLE_ARRAYS = {"[I": int, "[D": float, "[Z": bool, "[J": int, "[Ljava.lang.String;": str}
SCALARS = {"java.lang.Integer": int, "java.lang.Double": float, "java.lang.Boolean": bool, "java.lang.String": str, "java.lang.Long": int}


def java_double(v: float) -> str:
    """Double.toString: decimal notation for 1e-3 <= |v| < 1e7, computerized scientific notation otherwise."""
    v = float(v)
    if math.isnan(v):
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    r = repr(abs(v))
    mant, _, exp = r.partition("e")
    digits = mant.replace(".", "")
    point = mant.index(".") if "." in mant else len(mant)
    e10 = int(exp or 0) + point - 1 - (len(digits) - len(digits.lstrip("0")))  # exponent of the first significant digit
    digits = digits.strip("0") or "0"
    if 1e-3 <= abs(v) < 1e7:
        if e10 >= 0:
            ip, fp = digits[:e10 + 1].ljust(e10 + 1, "0"), digits[e10 + 1:]
            s = ip + "." + (fp or "0")
        else:
            s = "0." + "0" * (-e10 - 1) + digits
    else:
        s = digits[0] + "." + (digits[1:] or "0") + "E" + str(e10)
    return ("-" if v < 0 else "") + s


def _esc(s: str) -> str:
    return s.replace("&", "&amp;").replace("<", "&lt;").replace(">", "&gt;").replace('"', "&quot;").replace("'", "&apos;")


def _unesc(s: str) -> str:
    return s.replace("&lt;", "<").replace("&gt;", ">").replace("&quot;", '"').replace("&apos;", "'").replace("&amp;", "&")


class Storable:
    """An object with its own toXML() (className + body)."""

    def __init__(self, class_name: str, body: str):
        self.class_name, self.body = class_name, body


def _simple(v) -> str:
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    if isinstance(v, (float, np.floating)):
        return java_double(v)
    return _esc(str(v))


def append_object(tag: str, v, class_name: str | None = None, attrs: str | None = None, write_class: bool = True) -> str:
    """XMLParser.appendObjectWithTagsAndAttributes (:233-271)."""
    head = "<" + tag + ("" if attrs is None else " " + attrs) + ">"
    tail = "</" + tag + ">"
    if v is None:
        return head + "null" + tail
    out = [head]
    if isinstance(v, Storable):
        if write_class:
            out.append("<className>" + v.class_name + "</className>")
        out.append(v.body)
    elif isinstance(v, (list, tuple, np.ndarray)):
        assert class_name is not None and class_name.startswith("["), "arrays need their Java class name"
        if write_class:
            out.append("<className>" + class_name + "</className>")
        out.append("<length>" + str(len(v)) + "</length>")
        comp = class_name[1:]
        simple = class_name in SIMPLE_ARRAYS
        for i, x in enumerate(v):
            if simple:
                out.append(append_object("pos", x, None, f'val="{i}"', False))
            elif comp.startswith("["):
                out.append(append_object("pos", x, comp, f'val="{i}"', True))
            else:
                out.append(append_object("pos", x, None, f'val="{i}"', True))
    else:
        if write_class:
            cn = class_name or ("java.lang.Boolean" if isinstance(v, (bool, np.bool_)) else "java.lang.Integer" if isinstance(v, (int, np.integer))
                                else "java.lang.Double" if isinstance(v, (float, np.floating)) else "java.lang.String")
            out.append("<className>" + cn + "</className>")
        out.append(_simple(v))
    out.append(tail)
    return "".join(out)


def add_tags(body: str, tag: str) -> str:
    """XMLParser.addTags (:116-147)."""
    return "<" + tag + ">" + ("\n" if body else "") + body + "</" + tag + ">\n"


# ---- reading ---------------------------------------------------------------------------------------------------------------
class Node:
    __slots__ = ("tag", "attrs", "children", "text")

    def __init__(self, tag, attrs):
        self.tag, self.attrs, self.children, self.text = tag, attrs, [], ""

    def child(self, tag, **attrs):
        for c in self.children:
            if c.tag == tag and all(c.attrs.get(k) == str(v) for k, v in attrs.items()):
                return c
        return None

    def find(self, tag):
        """first element with this tag in document order (XMLParser.extractForTag searches the same way)"""
        for c in self.children:
            if c.tag == tag:
                return c
            r = c.find(tag)
            if r is not None:
                return r
        return None


_TOKEN = re.compile(r"<(/?)([\w$.:\-]+)((?:\s+[\w:\-]+\s*=\s*\"[^\"]*\")*)\s*(/?)>")


def parse(xml: str) -> Node:
    root = Node("#root", {})
    stack = [root]
    pos = 0
    for m in _TOKEN.finditer(xml):
        stack[-1].text += xml[pos:m.start()]
        pos = m.end()
        closing, tag, attrs, selfclose = m.group(1), m.group(2), m.group(3), m.group(4)
        if closing:
            if stack[-1].tag != tag:
                raise ValueError(f"malformed XML: </{tag}> closes <{stack[-1].tag}>")
            stack.pop()
        else:
            n = Node(tag, dict(re.findall(r"([\w:\-]+)\s*=\s*\"([^\"]*)\"", attrs)))
            stack[-1].children.append(n)
            if not selfclose:
                stack.append(n)
    if len(stack) != 1:
        raise ValueError("malformed XML: unclosed <" + stack[-1].tag + ">")
    return root


def _cast(t, s: str):
    s = s.strip()
    if t is bool:
        return s == "true"
    if t is float:
        return float(s.replace("Infinity", "inf").replace("NaN", "nan"))
    if t is int:
        return int(s)
    return _unesc(s)


def to_object(node: Node, class_hint: str | None = None):
    """XMLParser.extractObjectForTags (:687-785): null, arrays, simple types; Storables come back as (className, Node)."""
    if node is None:
        return None
    if not node.children and node.text.strip() == "null":
        return None
    cn = node.child("className")
    cls = cn.text.strip() if cn is not None else class_hint
    if cls is not None and cls.startswith("["):
        n = int(node.child("length").text)
        comp = cls[1:]
        out = []
        for i in range(n):
            p = node.child("pos", val=i)
            if cls in SIMPLE_ARRAYS:
                out.append(_cast(SIMPLE_ARRAYS[cls], p.text))
            else:
                out.append(to_object(p, comp if comp.startswith("[") else None))
        return out
    if cls in SCALARS:
        return _cast(SCALARS[cls], node.text)
    if cls is None:
        return _unesc(node.text.strip())
    return (cls, node)


# ---- the HMM document ------------------------------------------------------------------------------------------------------
def _dna_container_xml() -> Storable:
    alph = append_object("pos", Storable("de.jstacs.data.alphabets.DNAAlphabet", ""), None, 'val="0"')  # a Singleton: class name only
    body = "<Alphabets><className>[Lde.jstacs.data.alphabets.Alphabet;</className><length>1</length>" + alph + "</Alphabets>"
    return Storable("de.jstacs.data.AlphabetContainer", add_tags(body, "AlphabetContainer"))


def hmm_to_xml(hmm) -> str:
    """AbstractHMM.toXML (:394-415) of the Python mirror."""
    from .hmm import (DifferentiableHigherOrderHMM, HigherOrderTransition, IterationCondition, SilentEmission, TransitionElement,
                      ViterbiParameterSet)

    tr = hmm.transition
    diff = isinstance(tr, HigherOrderTransition)
    elems = []
    for t in tr.transitions:
        body = (append_object("context", t.context, "[I") + append_object("states", t.states, "[I") +
                append_object("hyperparameters", t.hyperParameters, "[D") + append_object("parameters", t.parameters, "[D") +
                append_object("weight", None) + append_object("norm", bool(t.norm)))
        cls = HMM_PKG + ".transitions.elements." + ("TransitionElement" if isinstance(t, TransitionElement) else "BasicTransitionElement")
        elems.append(Storable(cls, add_tags(body, "TRANSITION_ELEMENT")))
    arr_cls = "[L" + HMM_PKG + (".transitions.elements.TransitionElement;" if diff else ".transitions.BasicHigherOrderTransition$AbstractTransitionElement;")
    tbody = (append_object("transitions", elems, arr_cls) + append_object("isSilent", [bool(b) for b in tr.isSilent], "[Z") +
             append_object("transIndex", list(tr.transIndex), "[I"))
    if diff:
        tbody += append_object("offset", 0)
    ttag = "HigherOrderTransition" if diff else "BasicHigherOrderTransition"
    transition = Storable(HMM_PKG + ".transitions." + ttag, add_tags(tbody, ttag))

    ems = []
    for e in hmm.emission:
        if isinstance(e, SilentEmission):
            ems.append(Storable(HMM_PKG + ".states.emissions.SilentEmission", ""))  # SilentEmission.toXML() returns null
            continue
        rows = [list(map(float, r)) for r in e.params]
        zeros = [[0.0] * len(r) for r in rows]
        body = (append_object("alphabetContainer", _dna_container_xml()) + append_object("params", rows, "[[D") +
                append_object("norm", bool(e.norm)) + append_object("offset", 0) +
                append_object("hyperParams", [list(map(float, r)) for r in e.hyperParams], "[[D") +
                append_object("initHyperParams", [list(map(float, r)) for r in e.initHyperParams], "[[D") +
                append_object("statistic", zeros, "[[D") + append_object("ess", [float(np.sum(r)) for r in e.hyperParams], "[D") +
                append_object("shape", None) + append_object("linear", False) + append_object("hasParameters", False))
        kind = "DiscreteEmission" if type(e).__name__ == "DiscreteEmission" else "HigherOrderDiscreteEmission"
        if kind == "HigherOrderDiscreteEmission":
            body += append_object("order", int(e.order))
        ems.append(Storable(HMM_PKG + ".states.emissions.discrete." + kind, add_tags(body, "ConditionalDiscreteEmission")))

    tp = hmm.trainingParameter
    tc = tp.getTerminationCondition()
    tc_cls = "de.jstacs.algorithms.optimization.termination." + type(tc).__name__
    tc_val = tc.maxIter if isinstance(tc, IterationCondition) else tc.eps
    pset = add_tags(add_tags(append_object("numberOfStarts", tp.getNumberOfStarts()) + append_object("terminationCondition", tc_cls) +
                             append_object("terminationValue", tc_val) + append_object("threads", tp.getNumberOfThreads()), "set"), "parameterSet")
    tp_cls = HMM_PKG + ".training." + ("ViterbiParameterSet" if isinstance(tp, ViterbiParameterSet) else "BaumWelchParameterSet")
    n = len(hmm.name)
    doc = (append_object("trainingParameter", Storable(tp_cls, pset)) + "\n" + append_object("transition", transition) + "\n" +
           append_object("name", list(hmm.name), "[Ljava.lang.String;") + "\n" +
           append_object("filter", [None] * n, "[L" + HMM_PKG + ".states.filter.Filter;") + "\n" +
           append_object("emissionIdx", list(hmm.emissionIdx), "[I") + "\n" + append_object("strand", [True] * n, "[Z") + "\n" +
           append_object("emission", ems, "[L" + HMM_PKG + ".states.emissions.Emission;") + "\n" +
           append_object("statesGroups", [list(g) for g in getattr(hmm, "statesGroups", [])], "[[I") + "\n" +
           append_object("skipInit", bool(hmm.skipInit)) + append_object("type", getattr(hmm, "type", None)))
    if isinstance(hmm, DifferentiableHigherOrderHMM):
        doc += append_object("ess", float(hmm.ess))
    return add_tags(doc, "HigherOrderHMM")


def hmm_from_xml(xml: str):
    """AbstractHMM.fromXML (:424-465): rebuilds the Python mirror from a document written by hmm_to_xml or by Jstacs."""
    from . import hmm as H
    from .data import DNA

    root = parse(xml).find("HigherOrderHMM")
    if root is None:
        raise ValueError("no <HigherOrderHMM> element")
    # training parameters: values by tag search (Jstacs writes its generic parameter framework here), defaults otherwise
    tpn = root.child("trainingParameter")
    tp_cls = (tpn.child("className").text if tpn is not None and tpn.child("className") is not None else "")
    starts, threads, tc = 1, 1, H.IterationCondition(100)
    if tpn is not None:
        v = tpn.find("numberOfStarts")
        starts = int(to_object(v)) if v is not None else 1
        v = tpn.find("threads")
        threads = int(to_object(v)) if v is not None else 1
        c, val = tpn.find("terminationCondition"), tpn.find("terminationValue")
        if c is not None and val is not None:
            tc = (H.IterationCondition(int(to_object(val))) if str(to_object(c)).endswith("IterationCondition")
                  else H.SmallDifferenceOfFunctionEvaluationsCondition(float(to_object(val))))
    pars = (H.ViterbiParameterSet if tp_cls.endswith("ViterbiParameterSet") else H.BaumWelchParameterSet)(starts, tc, threads)
    # transition
    tcls, tnode = to_object(root.child("transition"))
    tnode = tnode.find("HigherOrderTransition") or tnode.find("BasicHigherOrderTransition")
    diff = tcls.endswith(".HigherOrderTransition")
    elements = []
    for ecls, en in to_object(tnode.child("transitions")):
        en = en.find("TRANSITION_ELEMENT")
        ctx, states = to_object(en.child("context")), to_object(en.child("states"))
        hyper, params = to_object(en.child("hyperparameters")), to_object(en.child("parameters"))
        norm = en.child("norm")
        ctor = H.TransitionElement if ecls.endswith(".TransitionElement") else H.BasicTransitionElement
        te = ctor(ctx, states, hyper, None, True if norm is None else bool(to_object(norm)))
        te.setLogProbabilities(params)
        elements.append(te)
    trans_index = to_object(tnode.child("transIndex"))
    names = to_object(root.child("name"))
    emission_idx = to_object(root.child("emissionIdx"))
    strand = to_object(root.child("strand"))
    emissions = []
    for ecls, en in to_object(root.child("emission")):
        if ecls.endswith("SilentEmission"):
            emissions.append(H.SilentEmission(DNA))
            continue
        if not (ecls.endswith(".HigherOrderDiscreteEmission") or ecls.endswith(".DiscreteEmission")):
            raise NotImplementedError(f"emission {ecls} is outside the native path")
        en = en.find("ConditionalDiscreteEmission")
        params, hyper = np.array(to_object(en.child("params"))), np.array(to_object(en.child("hyperParams")))
        init = en.child("initHyperParams")
        init = None if init is None or to_object(init) is None else np.array(to_object(init))
        norm = en.child("norm")
        order = int(to_object(en.child("order"))) if en.child("order") is not None else 0
        e = H.AbstractConditionalDiscreteEmission(DNA, hyper, init, True if norm is None else bool(to_object(norm)), order)
        e.__class__ = H.HigherOrderDiscreteEmission if ecls.endswith(".HigherOrderDiscreteEmission") else H.DiscreteEmission
        e.setLogProbabilities(params)
        emissions.append(e)
    ess = root.child("ess")
    kw = dict(emissionIdx=emission_idx, forward=strand, transIndex=trans_index)
    if diff and ess is not None:
        model = H.DifferentiableHigherOrderHMM(pars, names, emission_idx, strand, emissions, float(to_object(ess)), *elements, transIndex=trans_index)
    else:
        model = H.HigherOrderHMM(pars, names, emissions, *elements, **kw)
    # the constructors clone and re-initialise nothing, but precompute() recomputed logNorm from the parameters as the
    # reference's fromXML does (AbstractConditionalDiscreteEmission.java:548-557, BasicHigherOrderTransition.java:832-858)
    skip = root.child("skipInit")
    model.skipInit = bool(to_object(skip)) if skip is not None else False
    model.setOutputStream(None)
    return model
