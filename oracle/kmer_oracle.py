"""kmer_oracle.py -- CPU restatement of the reference assembler's graph build. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and the cpu_baseline leg of the k-mer benchmark may import this file; the product
path (ubu_b200/csrc/ubu_kmer.cu) never does. Every function cites the reference lines it follows; `M/` =
/root/reference/src/main/java/edu/unc/bioinf/ubu/.

Parity status: this path restates Java that EXISTS in the reference (unlike the alignment itself), and it is pinned against
the one fixture the reference holds for it: src/test/java/.../assembly/input1.fastq with k=3, minNodeFrequency=2 must
assemble to the contig GCTCCAG (expected output1.fasta next to it and the body of AssemblerTest.testBasicAssembly,
AssemblerTest.java:22-44; the test body is commented out in the reference, the fixture files are live). The reads of that
fixture are committed as tests/golden/assembler_input1.json.

Plain Python dicts/lists: meant for small cases (a few thousand reads).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple


class Node:
    """M/assembly/Node.java:10-26."""

    __slots__ = ("sequence", "count", "to_nodes", "from_nodes", "contributing_read", "has_multiple_unique_reads", "first")

    def __init__(self, sequence: str, first: Tuple[int, int]):
        self.sequence = sequence          # Node.java:31-33 (the k-mer; Sequence equality = equality of the bases, Sequence.java:83-86)
        self.count = 1                    # Node.java:13
        self.to_nodes: List["Node"] = []  # Node.java:17
        self.from_nodes: List["Node"] = []  # Node.java:18
        self.contributing_read: Optional[str] = None  # Node.java:24
        self.has_multiple_unique_reads = False        # Node.java:25
        self.first = first                # (read index, position) of creation: bookkeeping of this oracle, not in the Java

    def add_read_sequence(self, sequence: str) -> None:
        """Node.java:78-90."""
        if not self.has_multiple_unique_reads:
            if self.contributing_read is None:
                self.contributing_read = sequence
            elif self.contributing_read != sequence:
                self.has_multiple_unique_reads = True
                self.contributing_read = None

    def add_to_node(self, to: "Node") -> None:
        """Node.java:118-130: a new edge is appended to this.toNodes and to.fromNodes; a known one changes nothing."""
        if not any(n is to for n in self.to_nodes):      # findNode: equals() on the k-mer == identity inside one graph
            self.to_nodes.append(to)
            to.from_nodes.append(self)

    def is_root_node(self) -> bool:
        """Node.java:132-134."""
        return len(self.from_nodes) == 0

    def remove(self) -> None:
        """Node.java:200-208 with removeToNode / removeFromNode (:143-171): the other entries keep their order."""
        for to in self.to_nodes:
            to.from_nodes = [n for n in to.from_nodes if n is not self]
        for frm in self.from_nodes:
            frm.to_nodes = [n for n in frm.to_nodes if n is not self]


class Assembler:
    """The graph-building half of M/assembly/Assembler.java (one instance per region, ReAligner.java:328-349)."""

    def __init__(self, kmer_size: int, min_node_frequency: int):
        self.kmer_size = kmer_size                      # Assembler.java:165-167
        self.min_node_frequency = min_node_frequency    # Assembler.java:181-183
        self.nodes: Dict[str, Node] = {}                # Assembler.java:44 (insertion-ordered here; the Java HashMap is not)
        self.root_nodes: List[Node] = []                # Assembler.java:46
        self.filtered: Dict[str, Node] = {}             # nodes filterLowFrequencyNodes removed (kept for comparison only)

    def add_to_graph(self, sequence: str, read_index: int = 0) -> None:
        """Assembler.java:452-474."""
        prev = None
        k = self.kmer_size
        for i in range(0, len(sequence) - k + 1):
            kmer = sequence[i:i + k]
            node = self.nodes.get(kmer)
            if node is None:
                node = Node(kmer, (read_index, i))
                self.nodes[kmer] = node
            else:
                node.count += 1                          # incrementCount, Node.java:48-50
            node.add_read_sequence(sequence)
            if prev is not None:
                prev.add_to_node(node)
            prev = node

    def add_reads(self, reads: Sequence[str], first_index: int = 0) -> int:
        """The read loop of assembleContigs, Assembler.java:74-92: reads containing 'N' are skipped (the X0 > 1 test of the
        same lines needs a SAM tag and is the caller's). Returns the number skipped."""
        skipped = 0
        for j, r in enumerate(reads):
            if "N" in r:
                skipped += 1
            else:
                self.add_to_graph(r, first_index + j)
        return skipped

    def filter_low_frequency_nodes(self) -> None:
        """Assembler.java:201-227."""
        to_filter = [n for n in self.nodes.values() if n.count < self.min_node_frequency or not n.has_multiple_unique_reads]
        for n in to_filter:
            n.remove()
        for n in to_filter:
            self.filtered[n.sequence] = self.nodes.pop(n.sequence)

    def identify_root_nodes(self) -> None:
        """Assembler.java:270-276."""
        self.root_nodes = [n for n in self.nodes.values() if n.is_root_node()]

    # ---- the host-side contig walk, restated only so that the reference's fixture can pin the graph --------------------------
    def build_contigs(self, min_contig_length: int = 0, max_depth: int = 10000) -> List[str]:
        """Assembler.java:278-367 without the region-ratio test and the potential-contig cap (neither matters for the fixture).
        A contig = the first character of every node on the path + the whole k-mer of the terminal node (:300-303, :349)."""
        out: List[str] = []

        def walk(node: Node, visited: frozenset, contig: str, depth: int) -> None:
            if depth > max_depth:
                raise RecursionError("DepthExceededException")
            if node.sequence in visited:                 # terminated at a repeat: the node is not appended (:298-305)
                if len(contig) >= min_contig_length:
                    out.append(contig)
                return
            visited = visited | {node.sequence}
            if not node.to_nodes:
                contig = contig + node.sequence
                if len(contig) >= min_contig_length:
                    out.append(contig)
                return
            contig = contig + node.sequence[0]
            for to in node.to_nodes:
                walk(to, visited, contig, depth + 1)

        for root in self.root_nodes:
            walk(root, frozenset(), "", 1)
        return out


def build_graphs(reads: Sequence[str], regions: Optional[Sequence[int]], kmer_size: int, min_node_frequency: int):
    """Runs one Assembler per region over a batch (reads keep their batch index). Returns (nodes, skipped) where nodes is a list
    of dicts in the order (first_read, first_pos) -- the format tests compare the device table with:
        region, kmer, count, first_read, first_pos, multi, filtered, root, to (list of (region, kmer)), from (same)
    `to` / `from` are the UNFILTERED lists in creation order."""
    per_region: Dict[int, Assembler] = {}
    skipped = 0
    for j, r in enumerate(reads):
        reg = 0 if regions is None else int(regions[j])
        if reg == 0xFFFFFFFF:
            continue
        if "N" in r:
            skipped += 1
            continue
        a = per_region.get(reg)
        if a is None:
            a = per_region[reg] = Assembler(kmer_size, min_node_frequency)
        a.add_to_graph(r, j)
    rows = []
    for reg, a in per_region.items():
        before = {k: (list(n.to_nodes), list(n.from_nodes)) for k, n in a.nodes.items()}     # lists as addToGraph left them
        all_nodes = dict(a.nodes)
        a.filter_low_frequency_nodes()
        a.identify_root_nodes()
        roots = {n.sequence for n in a.root_nodes}
        for kmer, n in all_nodes.items():
            to, frm = before[kmer]
            rows.append({"region": reg, "kmer": kmer, "count": n.count, "first_read": n.first[0], "first_pos": n.first[1],
                         "multi": n.has_multiple_unique_reads, "filtered": kmer in a.filtered, "root": kmer in roots,
                         "to": [(reg, t.sequence) for t in to], "from": [(reg, f.sequence) for f in frm]})
    rows.sort(key=lambda d: (d["first_read"], d["first_pos"]))
    return rows, skipped
