"""Writes the small fixture networks under tests/golden/ (run once; outputs are committed).

  cancer.net      byte copy of the reference's only bundled data file is NOT made; instead the same network is
                  re-emitted from its CPT numbers in this repo's own writer (node order D,E,A,C,B as in
                  /root/reference/cancer.net:53-102, children before parents -> exercises the moralisation quirk A.6).
  sprinkler.net   the `example` network of Bayes/Examples.hs:217-233 (ids winter0 sprinkler1 wet2 rain3 road4 rr5).
  asia.net        Lauritzen-Spiegelhalter "Asia" with the canonical CPTs (the reference's asia.net is not bundled;
                  goldens at Bayes/Test/ReferencePatterns.hs:119-131).
  asia_rev.net    same network, nodes listed children-first (moralisation quirk active).
  sprinkler_gui.net  the sprinkler network again, written the way a Hugin / SamIam GUI export looks (the dialect of the
                  reference's bundled cancer.net: a `net { HR_... = "..."; }` header full of tool properties, node
                  sections with position / label / ID attributes, tab-indented nested `data = (( ... )( ... ));`
                  blocks, `potential ( A | )` for a root) -- content and property values are this repo's own; only
                  the FORMAT follows /root/reference/cancer.net:1-128.  Must load to exactly sprinkler.net's network.
"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from hbayes_b200.synth import SynthNet

HERE = os.path.dirname(os.path.abspath(__file__))

def net_from(names, dims, cpts):
    """cpts: {child: (parents, values in reference layout [child, parents...])}"""
    ids = {n: i for i, n in enumerate(names)}
    parents = [[ids[p] for p in cpts[n][0]] for n in names]
    values = [cpts[n][1] for n in names]
    return SynthNet(dims, parents, values, names)

cancer = net_from(["D", "E", "A", "C", "B"], [2] * 5, {
    "D": (["C", "B"], [.8, .8, .8, .05, .2, .2, .2, .95]),
    "E": (["C"], [.8, .6, .2, .4]),
    "A": ([], [.2, .8]),
    "C": (["A"], [.2, .05, .8, .95]),
    "B": (["A"], [.8, .2, .2, .8]),
})
sprinkler = net_from(["winter", "sprinkler", "wet", "rain", "road", "roadandrain"], [2] * 6, {
    "winter": ([], [0.4, 0.6]),
    "sprinkler": (["winter"], [0.25, 0.8, 0.75, 0.2]),
    "rain": (["winter"], [0.9, 0.2, 0.1, 0.8]),
    "wet": (["sprinkler", "rain"], [1, 0.2, 0.1, 0.05, 0, 0.8, 0.9, 0.95]),
    "road": (["rain"], [1, 0.3, 0, 0.7]),
    # logical roadandrain ((rain .==. True) .&. (road .==. True))  (BayesianNetwork.hs:161-171)
    "roadandrain": (["rain", "road"], [1, 1, 1, 0, 0, 0, 0, 1]),
})
# state 0 = "yes", state 1 = "no"
asia_cpts = {
    "A": ([], [0.01, 0.99]),
    "S": ([], [0.5, 0.5]),
    "T": (["A"], [0.05, 0.01, 0.95, 0.99]),
    "L": (["S"], [0.1, 0.01, 0.9, 0.99]),
    "B": (["S"], [0.6, 0.3, 0.4, 0.7]),
    "E": (["L", "T"], [1, 1, 1, 0, 0, 0, 0, 1]),
    "X": (["E"], [0.98, 0.05, 0.02, 0.95]),
    "D": (["E", "B"], [0.9, 0.7, 0.8, 0.1, 0.1, 0.3, 0.2, 0.9]),
}
asia = net_from(["A", "S", "T", "L", "B", "E", "X", "D"], [2] * 8, asia_cpts)
asia_rev = net_from(["X", "D", "E", "B", "L", "T", "S", "A"], [2] * 8, asia_cpts)

def gui_text(net):
    """the network in the GUI-export dialect (see the module docstring)"""
    props = [("HR_Compile_TriangMethod", "0"), ("HR_Monitor_GraphPrecision", "100"), ("HRUNTIME_Grid_GridSnap", "1"),
             ("HR_Propagate_AutoNormal", "1"), ("HR_Font_Size", "-12"), ("HR_Font_Name", "Helvetica"), ("HR_Grid_X", "10"), ("HR_Grid_Y", "10"),
             ("HRUNTIME_Compile_ApproxEpsilon", "0.00001"), ("HR_Groups_GroupNames", ""), ("HR_Color_DiscreteChance", "16")]
    out = ["net", "{"]
    for i, (k, v) in enumerate(props):
        out.append('\t%s = "%s";' % (k, v))
        if i == 4:
            out.append("\tnode_size = (100.0 40.0);")
    out += ["}", ""]
    for v, name in enumerate(net.names):
        out += ["node " + name, "{", '\tstates = ("true" "false" );', "\tposition = (%d %d);" % (100 + 130 * (v % 3), -90 * (v // 3)),
                '\texcludepolicy = "include whole CPT";', '\tismapvariable = "false";', '\tID = "%s";' % name,
                '\tlabel = "%s: node %d";' % (name, v), '\tdiagnosistype = "AUXILIARY";', "}"]

    def block(vals, dims):  # nested parentheses, one group per parent configuration, child fastest
        if len(dims) == 1:
            return "(\t" + "\t".join(repr(float(x)) for x in vals) + "\t)"
        step = len(vals) // dims[0]
        return "(" + "\n\t\t".join(block(vals[i * step:(i + 1) * step], dims[1:]) for i in range(dims[0])) + ")"

    import numpy as np
    for v, name in enumerate(net.names):
        pa = net.parents[v]
        # file layout: parents slowest, child fastest (Bayes/ImportExport/HuginNet.hs:139-147)
        table = np.asarray(net.values[v], float).reshape([net.dims[v]] + [net.dims[p] for p in pa])
        filed = np.moveaxis(table, 0, -1).reshape(-1)
        out += ["potential ( %s | %s)" % (name, "".join(net.names[p] + " " for p in pa)), "{",
                "\tdata = " + block(filed.tolist(), [net.dims[p] for p in pa] + [net.dims[v]]) + ";", "}"]
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    open(os.path.join(HERE, "sprinkler_gui.net"), "w").write(gui_text(sprinkler))
    print("wrote sprinkler_gui")
    for name, net in [("cancer", cancer), ("sprinkler", sprinkler), ("asia", asia), ("asia_rev", asia_rev)]:
        net.write_hugin(os.path.join(HERE, name + ".net"))
        print("wrote", name)
