"""hbayes_b200 -- host-side mirror of hbayes' exact-inference API over libhbayes_cuda (B200, sm_100a).

The names follow the reference (alpheccar/hbayes) so that its examples and tests read the same:

    importBayesianGraph   Bayes/ImportExport/HuginNet.hs:160-167
    createJunctionTree    Bayes/FactorElimination.hs:298-307
    changeEvidence        Bayes/FactorElimination/JTree.hs:454-466
    posterior             Bayes/FactorElimination.hs:314-327
    priorMarginal         Bayes/VariableElimination.hs:134-147
    posteriorMarginal     Bayes/VariableElimination.hs:116-130
    minDegreeOrder / minFillOrder   Bayes/VariableElimination.hs:225-236
    changeFactor          Bayes/Factor.hs:66-81, Bayes/FactorElimination/JTree.hs:107-115
    learnEM               Bayes/EM.hs:36-45
    mpe                   Bayes/VariableElimination.hs:101-114

plus the batched forms (one evidence row per query) that the GPU exists for.  All arithmetic happens in
libhbayes_cuda.so; this package only marshals arrays.  It never imports `oracle/`.
"""
from .api import (  # noqa: F401
    BayesianNetwork,
    Factor,
    JunctionTree,
    MostProbableExplanation,
    VariableElimination,
    changeEvidence,
    changeFactor,
    createJunctionTree,
    degreeOrder,
    factorProduct,
    factorProjectOut,
    factorProjectTo,
    host_alloc,
    nccl_unique_id,
    marginalizeOneVariable,
    importBayesianGraph,
    learnEM,
    minDegreeOrder,
    minFillOrder,
    mpe,
    nodeComparisonForTriangulation,
    numberOfAddedEdges,
    posterior,
    posteriorMarginal,
    priorMarginal,
)
from ._lib import HbError, LIB_PATH  # noqa: F401
