#!/usr/bin/env python
"""Stand-in for `samtools` in tests (samtools is not installed in the build image).

  fake_samtools.py faidx <fasta> <ctg:start-end>      -> FASTA record of the region (1-based, inclusive)
  fake_samtools.py mpileup ... -r <ctg:start-end> --min-BQ <k> ... <bam>
                                                     -> rows of <bam>.minbq<k>.mpileup inside the region
Both the unmodified reference (tests/golden/make_golden.py) and the drop-in are run against it.
"""
import sys


def read_fasta(path):
    seqs, name = {}, None
    for line in open(path):
        line = line.rstrip("\n")
        if line.startswith(">"):
            name = line[1:].split()[0]
            seqs[name] = []
        elif name:
            seqs[name].append(line)
    return {k: "".join(v) for k, v in seqs.items()}


def parse_region(region):
    ctg, span = region.rsplit(":", 1)
    a, b = span.split("-")
    return ctg, int(a), int(b)


def main(argv):
    if argv[0] == "faidx":
        ctg, a, b = parse_region(argv[2])
        seq = read_fasta(argv[1])[ctg][a - 1:b]
        print(">%s" % argv[2])
        for i in range(0, len(seq), 60):
            print(seq[i:i + 60])
        return 0
    if argv[0] == "mpileup":
        region = argv[argv.index("-r") + 1]
        min_bq = argv[argv.index("--min-BQ") + 1]
        ctg, a, b = parse_region(region)
        for line in open("%s.minbq%s.mpileup" % (argv[-1], min_bq)):
            cols = line.split("\t", 2)
            if cols[0] == ctg and a <= int(cols[1]) <= b:
                sys.stdout.write(line)
        return 0
    sys.stderr.write("fake_samtools: unsupported command %r\n" % argv[:1])
    return 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
