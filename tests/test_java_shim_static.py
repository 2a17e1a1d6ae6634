"""Static checks of the Java drop-in (java/de/jstacs/...) — the image has no JDK, so the shim cannot be compiled here.

Two groups:
  * binding vs header (always run): every Panama downcall descriptor in JstacsHmmNative.java matches the prototype in
    include/jstacs_hmm.h; the jh_model_desc / jh_train_opts layouts used by GpuHmmEngine.java match the C structs;
  * shim vs reference sources (skipped when /root/reference is absent, e.g. on the GPU box): every imported Jstacs class
    exists, every inherited member the shim touches is declared in the reference with the type and visibility the shim
    assumes, no shim method overrides a `final` method, every @Override has a target in the superclass chain, and the
    semantics the shim depends on (statesGroups is an empty array — never null — for "no groups"; IntList rejects size 0;
    forward/filter are never null) still hold in the reference sources.
"""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java")
REF = "/root/reference"
HMM = "de/jstacs/sequenceScores/statisticalModels/trainable/hmm"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "de", "jstacs")), reason="reference tree not present")


def java_sources():
    out = {}
    for d, _, files in os.walk(JAVA):
        for f in files:
            if f.endswith(".java"):
                p = os.path.join(d, f)
                out[os.path.relpath(p, JAVA)] = open(p, encoding="utf-8").read()
    return out


def strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def ref(path):
    return open(os.path.join(REF, path), encoding="utf-8", errors="replace").read()


# ---------------------------------------------------------------------------------------------------------------------
# binding vs header
def header_prototypes():
    h = strip_comments(open(os.path.join(ROOT, "include", "jstacs_hmm.h")).read())
    protos = {}
    for m in re.finditer(r"\b(int|int64_t|void|const char \*)\s*(jh_\w+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        kinds = []
        if args != "void":
            for a in args.split(","):
                a = a.strip()
                if "*" in a or "[" in a:
                    kinds.append("P")
                elif re.match(r"(const )?(int64_t)\b", a):
                    kinds.append("J")
                elif re.match(r"(const )?(int|int32_t)\b", a):
                    kinds.append("I")
                elif re.match(r"(const )?double\b", a):
                    kinds.append("D")
                else:
                    raise AssertionError(f"unmapped C type in {name}: {a}")
        protos[name] = ({"int": "I", "int64_t": "J", "void": "V", "const char *": "P"}[ret], kinds)
    return protos


def test_downcall_descriptors_match_the_header():
    src = strip_comments(java_sources()["de/jstacs/gpu/JstacsHmmNative.java"])
    protos = header_prototypes()
    seen = 0
    for m in re.finditer(r'h\(\s*"(jh_\w+)"\s*,\s*FunctionDescriptor\.(ofVoid|of)\(([^)]*)\)\s*\)', src):
        name, kind, args = m.group(1), m.group(2), [a.strip() for a in m.group(3).split(",") if a.strip()]
        assert name in protos, f"{name} is bound in Java but not declared in include/jstacs_hmm.h"
        ret, kinds = protos[name]
        if kind == "ofVoid":
            assert ret == "V", name
            got = args
        else:
            assert args[0] == ret, f"{name}: return {args[0]} != {ret}"
            got = args[1:]
        assert got == kinds, f"{name}: Java descriptor {got} != header {kinds}"
        seen += 1
    assert seen >= 25
    # every entry point the engine invokes is bound
    eng = java_sources()[f"{HMM}/models/GpuHmmEngine.java"]
    for name in set(re.findall(r"JstacsHmmNative\.(jh_\w+)\.invokeExact", eng)):
        assert re.search(rf'MethodHandle {name}\b', src), f"{name} is invoked but not bound"


def test_abi_version_constant_matches_the_header():
    h = open(os.path.join(ROOT, "include", "jstacs_hmm.h")).read()
    v = int(re.search(r"#define JH_ABI_VERSION (\d+)", h).group(1))
    src = java_sources()["de/jstacs/gpu/JstacsHmmNative.java"]
    assert int(re.search(r"JH_ABI_VERSION = (\d+)", src).group(1)) == v
    for name, val in re.findall(r"(JH_(?:OK|ERR_\w+)) = (-?\d+)", h):
        m = re.search(rf"\b{name} = (-?\d+)", src)
        assert m and int(m.group(1)) == int(val), name


def test_model_desc_layout_matches_the_c_struct():
    h = strip_comments(open(os.path.join(ROOT, "include", "jstacs_hmm.h")).read())
    body = re.search(r"typedef struct jh_model_desc \{(.*?)\} jh_model_desc;", h, flags=re.S).group(1)
    c_fields = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        name = re.search(r"(\w+)$", decl).group(1)
        c_fields.append((name, "P" if "*" in decl else "I"))
    eng = java_sources()[f"{HMM}/models/GpuHmmEngine.java"]
    lay = re.search(r"DESC = MemoryLayout\.structLayout\((.*?)\);", eng, flags=re.S).group(1)
    j_fields, offset = [], 0
    for m in re.finditer(r'ValueLayout\.(JAVA_INT|ADDRESS)\.withName\( "(\w+)" \)|MemoryLayout\.paddingLayout\( (\d+) \)', lay):
        if m.group(3):
            offset += int(m.group(3))
            continue
        size = 4 if m.group(1) == "JAVA_INT" else 8
        assert offset % size == 0, f"{m.group(2)} is misaligned at {offset}"
        j_fields.append((m.group(2), "I" if size == 4 else "P", offset))
        offset += size
    assert [(n, k) for n, k, _ in j_fields] == c_fields
    # C offsets with natural alignment
    off = 0
    for (name, kind), (_, _, joff) in zip(c_fields, j_fields):
        size = 4 if kind == "I" else 8
        off = (off + size - 1) // size * size
        assert off == joff, f"{name}: C offset {off} != Java offset {joff}"
        off += size
    # every field is filled
    for name, _ in c_fields:
        assert f'put( d, "{name}"' in eng, f"jh_model_desc.{name} is never set"
    # jh_train_opts {int32 mode, int32 term_kind, int32 max_iter, double eps}: offsets 0,4,8,16, size 24
    assert re.search(r"opts = arena\.allocate\( 24, 8 \)", eng)
    for off_, what in ((0, "mode"), (4, "kind"), (8, "maxIter")):
        assert re.search(rf"opts\.set\( ValueLayout\.JAVA_INT, {off_}, {what} \)", eng)
    assert re.search(r"opts\.set\( ValueLayout\.JAVA_DOUBLE, 16, eps \)", eng)


# ---------------------------------------------------------------------------------------------------------------------
# shim vs reference
@needs_ref
def test_every_imported_jstacs_class_exists():
    for path, src in java_sources().items():
        for imp in re.findall(r"^import (de\.jstacs[\w.]*);", src, flags=re.M):
            parts = imp.split(".")
            # nested classes: walk back until a file exists
            for cut in range(len(parts), 2, -1):
                rel = "/".join(parts[:cut]) + ".java"
                if os.path.exists(os.path.join(REF, rel)) or os.path.exists(os.path.join(JAVA, rel)):
                    nested = parts[cut:]
                    if nested:
                        text = ref(rel) if os.path.exists(os.path.join(REF, rel)) else open(os.path.join(JAVA, rel)).read()
                        for n in nested:
                            assert re.search(rf"\b(class|interface|enum)\s+{n}\b", text), f"{imp}: no nested type {n} in {rel}"
                    break
            else:
                raise AssertionError(f"{path}: import {imp} does not resolve in the reference tree")


CHAIN = {
    "GpuHigherOrderHMM": [f"{HMM}/models/DifferentiableHigherOrderHMM.java", f"{HMM}/models/HigherOrderHMM.java", f"{HMM}/AbstractHMM.java",
                          "de/jstacs/sequenceScores/statisticalModels/trainable/AbstractTrainableStatisticalModel.java"],
    "GpuBasicHigherOrderHMM": [f"{HMM}/models/HigherOrderHMM.java", f"{HMM}/AbstractHMM.java",
                               "de/jstacs/sequenceScores/statisticalModels/trainable/AbstractTrainableStatisticalModel.java"],
}

# inherited members the shim classes read: name -> (declaring file, declaration regex)
MEMBERS = {
    "states": (f"{HMM}/AbstractHMM.java", r"protected State\[\] states;"),
    "emission": (f"{HMM}/AbstractHMM.java", r"protected Emission\[\] emission;"),
    "emissionIdx": (f"{HMM}/AbstractHMM.java", r"protected int\[\] emissionIdx;"),
    "forward": (f"{HMM}/AbstractHMM.java", r"protected boolean\[\] forward;"),
    "filter": (f"{HMM}/AbstractHMM.java", r"protected Filter\[\] filter;"),
    "transition": (f"{HMM}/AbstractHMM.java", r"protected Transition transition;"),
    "finalState": (f"{HMM}/AbstractHMM.java", r"protected boolean\[\] finalState;"),
    "statesGroups": (f"{HMM}/AbstractHMM.java", r"protected int\[\]\[\] statesGroups\b"),
    "trainingParameter": (f"{HMM}/AbstractHMM.java", r"protected HMMTrainingParameterSet trainingParameter;"),
    "sostream": (f"{HMM}/AbstractHMM.java", r"protected SafeOutputStream sostream;"),
    "skipInit": (f"{HMM}/models/HigherOrderHMM.java", r"protected boolean skipInit;"),
    "type": (f"{HMM}/models/HigherOrderHMM.java", r"protected String type;"),
    "initialize": (f"{HMM}/models/HigherOrderHMM.java", r"protected void initialize\( DataSet data, double\[\] weight \) throws Exception"),
    "createStates": (f"{HMM}/models/HigherOrderHMM.java", r"protected void createStates\(\)"),
}


@needs_ref
@pytest.mark.parametrize("cls", list(CHAIN))
def test_shim_class_against_its_superclass_chain(cls):
    src = strip_comments(java_sources()[f"{HMM}/models/{cls}.java"])
    chain = [strip_comments(ref(p)) for p in CHAIN[cls]]
    # the declared superclass is the head of the chain
    sup = re.search(rf"class {cls} extends (\w+)", src).group(1)
    assert re.search(rf"class {sup}\b", chain[0])
    # Host accessors return inherited members that exist with the assumed type
    used = set(re.findall(r"\{ return (\w+); \}", src)) | {"initialize", "createStates"}
    for name in used:
        assert name in MEMBERS, f"{cls} reads '{name}', which the static check does not know"
        path, decl = MEMBERS[name]
        assert path in CHAIN[cls], f"{name} is declared in {path}, outside the superclass chain of {cls}"
        assert re.search(decl, ref(path)), f"{name}: declaration '{decl}' not found in {path}"
    # no method of the shim collides with a final method of a superclass
    finals = set()
    for text in chain:
        finals |= set(re.findall(r"\bfinal\s+[\w<>\[\],.]+\s+(\w+)\s*\(", text))
    declared = set(re.findall(r"public\s+[\w<>\[\],. ]+?\s+(\w+)\s*\([^)]*\)\s*(?:throws [\w., ]+)?\s*\{", src))
    assert not (declared & finals), f"{cls} redeclares final methods {declared & finals}"
    # every @Override has a same-named target with the same number of parameters (close() comes from AutoCloseable)
    for m in re.finditer(r"@Override\s+(?:@\w+\([^)]*\)\s+)?public\s+[\w<>\[\],. ]+?\s+(\w+)\s*\(([^)]*)\)", src):
        name, n_par = m.group(1), len([a for a in m.group(2).split(",") if a.strip()])
        if name == "close":
            continue
        ok = False
        for text in chain:
            for t in re.finditer(rf"(?:public|protected)\s+[\w<>\[\],. ]+?\s+{name}\s*\(([^)]*)\)", text):
                if len([a for a in t.group(1).split(",") if a.strip()]) == n_par:
                    ok = True
        assert ok, f"{cls}.{name}/{n_par} overrides nothing in the superclass chain"
    # the constructors it forwards to exist with the same number of parameters
    for m in re.finditer(r"super\(([^)]*)\);", src):
        n_par = len([a for a in m.group(1).split(",") if a.strip()])
        ctor = [len([a for a in t.group(1).split(",") if a.strip()]) for t in re.finditer(rf"public {sup}\s*\(([^)]*)\)", chain[0])]
        assert n_par in ctor, f"{cls}: no {sup} constructor with {n_par} parameters (have {ctor})"


@needs_ref
def test_reference_semantics_the_shim_relies_on():
    ahmm = ref(f"{HMM}/AbstractHMM.java")
    # "no groups" is an EMPTY array, never null: the shim must test the length (round-1 bug: tested != null and always threw)
    assert re.search(r"if\( statesGroups==null \|\| statesGroups\.length==0 \) \{\s*this\.statesGroups = new int\[0\]\[\];", ahmm)
    eng = java_sources()[f"{HMM}/models/GpuHmmEngine.java"]
    assert re.search(r"groups != null && groups\.length > 0", eng)
    assert "fillPreComputedContext(statesGroups);" in ahmm  # the constructor always runs it
    # forward / filter are never null after the constructor
    assert "this.forward = new boolean[n];" in ahmm and "this.filter = new Filter[n];" in ahmm
    # IntList refuses a non-positive initial size: the shim clamps to 1
    il = ref("de/jstacs/utils/IntList.java")
    assert re.search(r"if\( size <= 0 \) \{\s*throw new IllegalArgumentException", il)
    assert not re.search(r"new IntList\( \(int\) L \)", eng)
    # termination conditions expose their value as parameter 0 of the current parameter set (Integer / Double)
    ic = ref("de/jstacs/algorithms/optimization/termination/IterationCondition.java")
    assert "DataType.INT" in ic and "getParameterAt( 0 ).setValue( maxIter )" in ic
    sd = ref("de/jstacs/algorithms/optimization/termination/SmallDifferenceOfFunctionEvaluationsCondition.java")
    assert "DataType.DOUBLE" in sd and "(Double) parameter.getParameterAt(0).getValue()" in sd
    assert "public AbstractTerminationConditionParameterSet getCurrentParameterSet() throws Exception" in ref(
        "de/jstacs/algorithms/optimization/termination/AbstractTerminationCondition.java")
    # Time / SafeOutputStream / Pair / State API
    assert "public static Time getTimeInstance( OutputStream out ) throws IOException" in ref("de/jstacs/utils/Time.java")
    assert "public class SafeOutputStream extends OutputStream" in ref("de/jstacs/utils/SafeOutputStream.java")
    assert "public Pair( E1 element1, E2 element2 )" in ref("de/jstacs/utils/Pair.java")
    assert "public boolean isSilent();" in ref(f"{HMM}/states/State.java")
    assert "public int getNumberOfStarts()" in ref(f"{HMM}/training/HMMTrainingParameterSet.java")
    # getAlphabetContainer / getLength are final in the model base class: the Host interface is satisfied by inheritance
    base = ref("de/jstacs/sequenceScores/statisticalModels/trainable/AbstractTrainableStatisticalModel.java")
    assert "public final AlphabetContainer getAlphabetContainer()" in base and "public final int getLength()" in base
    # allowed states groups: what GpuHmmEngine.attachGroups reads (HigherOrderHMM.doOneStep :810-820 does the same)
    assert "public SequenceAnnotation getSequenceAnnotationByType(String type, int idx)" in ref("de/jstacs/data/sequences/Sequence.java")
    assert "public class SequenceAnnotation extends ResultSet" in ref("de/jstacs/data/sequences/annotation/SequenceAnnotation.java")
    assert "public Result getResultAt(int index)" in ref("de/jstacs/results/ResultSet.java")
    assert "public Storable getResultInstance()" in ref("de/jstacs/results/StorableResult.java")
    assert re.search(r"public static class AllowedStatesGroups implements Storable \{\s*public int\[\] groups;", ahmm)
    assert "protected String type;" in ref(f"{HMM}/models/HigherOrderHMM.java")
    assert "attachGroups( arena, ds, data );" in eng and "jh_dataset_set_groups.invokeExact( handles.dataset, MemorySegment.NULL )" in eng
    # numerical training stays on the reference code path
    assert "if( trainingParameter instanceof NumericalHMMTrainingParameterSet )" in ref(f"{HMM}/models/DifferentiableHigherOrderHMM.java")


# protected fields the package-local exporters read / write: (exporter file, reference file, [declarations])
EXPORTERS = [
    (f"{HMM}/transitions/GpuTransitionExport.java", f"{HMM}/transitions/BasicHigherOrderTransition.java",
     {"transitions": r"protected AbstractTransitionElement\[\] transitions;", "maximalMarkovOrder": r"protected int maximalMarkovOrder;",
      "lookup": r"protected int\[\]\[\]\[\] lookup;", "transIndex": r"protected int\[\] transIndex;",
      "parameters": r"protected double\[\] parameters;", "hyperParameters": r"protected double\[\] hyperParameters;",
      "statistic": r"protected double\[\] statistic;", "logNorm": r"protected double logNorm;", "descendants": r"protected int\[\] descendants;",
      "states": r"public int\[\] states;", "getLastContextState": r"public int getLastContextState\(\)", "getNumberOfChildren": r"public (final )?int getNumberOfChildren\(\)"}),
    (f"{HMM}/states/emissions/discrete/GpuEmissionExport.java", f"{HMM}/states/emissions/discrete/AbstractConditionalDiscreteEmission.java",
     {"params": r"protected double\[\]\[\] params;", "probs": r"protected double\[\]\[\] probs;", "hyperParams": r"protected double\[\]\[\] hyperParams;",
      "statistic": r"protected double\[\]\[\] statistic;", "logNorm": r"protected double\[\] logNorm;"}),
]


@needs_ref
@pytest.mark.parametrize("exporter,reference,members", EXPORTERS, ids=["transition", "emission"])
def test_package_local_exporters_against_the_reference(exporter, reference, members):
    src = strip_comments(java_sources()[exporter])
    text = ref(reference)
    # the exporter lives in the package of the class whose protected members it reads
    pkg = re.search(r"^package ([\w.]+);", java_sources()[exporter], flags=re.M).group(1)
    assert pkg.replace(".", "/") == os.path.dirname(reference)
    for name, decl in members.items():
        assert re.search(rf"\b{name}\b", src), f"{exporter} no longer uses {name}: drop it from the table"
        assert re.search(decl, text), f"{name}: '{decl}' not found in {reference}"
    if "Emission" in exporter:
        ho = ref(f"{HMM}/states/emissions/discrete/HigherOrderDiscreteEmission.java")
        assert "protected int order;" in ho
        assert "public class DiscreteEmission extends AbstractConditionalDiscreteEmission" in ref(f"{HMM}/states/emissions/discrete/DiscreteEmission.java")
