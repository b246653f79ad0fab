"""VCF text for the records of the callers.

The reference writes VCF through ADAM (`toVariantContext.coalesce(1, shuffle = true).saveAsVcf`, Common.scala:290-293) and no
test of it pins the text, so byte parity with a JVM run is not defined; this writer follows the conventions SURVEY
Appendix A lists — one line per (contig, start, ref, alt), POS = start + 1, per-sample GT from the bdg-formats
GenotypeAllele pair (Ref -> 0, Alt -> 1, OtherAlt -> 2 when the line's other alternate is known, NoCall -> .), GQ =
genotypeQuality, DP = readDepth, AD = reference, alternate depths where the record carries them — and sorts the lines, so
that "identical VCF" (identical sets of fields) becomes identical files.
"""
from typing import Dict, Iterable, List, Sequence

from . import abi

_GT = {abi.GT_REF: "0", abi.GT_ALT: "1", abi.GT_OTHER_ALT: "2", abi.GT_NO_CALL: "."}


def _line_fields(g: dict) -> Dict[str, str]:
    """FORMAT fields of one record: threshold records carry only the genotype; called / somatic records also carry the
    fields AlleleConversions.scala:30-62 sets (genotypeQuality, readDepth, reference / alternate read depth)."""
    f = {"GT": "/".join(_GT[a] for a in g["gt"])}
    ev = g.get("evidence") or g.get("tumor")
    if ev is not None:
        f["GQ"] = str(g["phred"])
        f["DP"] = str(ev["read_depth"])
        f["AD"] = f"{ev['read_depth'] - ev['allele_read_depth']},{ev['allele_read_depth']}"
    return f


def vcf_lines(genotypes: Iterable[dict], contig_names: Sequence[str], sample_names: Sequence[str]) -> List[str]:
    """Body lines (no header), sorted by (contig index, position, ref, alt).  `genotypes` = Result.genotypes()."""
    rows: Dict[tuple, Dict[int, Dict[str, str]]] = {}
    for g in genotypes:
        # An empty alternate (a locus INSIDE a deletion: the reference's MidDeletion allele, ref = the deleted base, alt = "")
        # is spelt with VCF 4.2's own allele for "missing due to an upstream deletion": "*".  (The deletion itself is the
        # record at its anchor locus, REF = anchor + deleted bases.)  The reference does not pin the VCF text.
        alt = g["alt"] if g["alt"] else "*" if g["ref"] else "."
        if g["alt"] == "<ALT>":  # the symbolic allele of emit_ref / emit_no_call records
            alt = "."
        key = (g["contig"], g["start"], g["ref"] or "N", alt)
        rows.setdefault(key, {})[g["sample"]] = _line_fields(g)
    out = []
    for (contig, start, ref, alt) in sorted(rows):
        per_sample = rows[(contig, start, ref, alt)]
        keys = [k for k in ("GT", "GQ", "DP", "AD") if any(k in f for f in per_sample.values())]
        cols = [contig_names[contig], str(start + 1), ".", ref, alt, ".", ".", ".", ":".join(keys)]
        for s in range(len(sample_names)):
            f = per_sample.get(s)
            cols.append(":".join(f.get(k, ".") for k in keys) if f else ":".join("./." if k == "GT" else "." for k in keys))
        out.append("\t".join(cols))
    return out


def write_vcf(path: str, genotypes: Iterable[dict], contig_names: Sequence[str], sample_names: Sequence[str],
              contig_lengths: Sequence[int] = ()) -> int:
    """Writes a VCF 4.2 file; returns the number of body lines."""
    lines = vcf_lines(genotypes, contig_names, sample_names)
    with open(path, "w") as fh:
        fh.write("##fileformat=VCFv4.2\n##source=guacamole_b200\n")
        for i, name in enumerate(contig_names):
            ln = f",length={int(contig_lengths[i])}" if len(contig_lengths) > i else ""
            fh.write(f"##contig=<ID={name}{ln}>\n")
        fh.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        fh.write('##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Phred-scaled genotype quality">\n')
        fh.write('##FORMAT=<ID=DP,Number=1,Type=Integer,Description="Read depth">\n')
        fh.write('##FORMAT=<ID=AD,Number=R,Type=Integer,Description="Reference, alternate read depth">\n')
        fh.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(sample_names) + "\n")
        for ln in lines:
            fh.write(ln + "\n")
    return len(lines)
