"""The index-building subcommands of graph_kmer_index/command_line_interface.py with their flags unchanged
(SURVEY.md 8f-1): ``make_from_flat`` (cli:156-174, 267-276), ``make_reverse`` (cli:177-181, 278-281),
``make_reference_kmer_index`` (cli:184-193, 290-297) and ``add_reverse_complements`` (cli:656-667).  Everything numeric
runs in libgki.so; the files written are the reference's npz layouts.  The rest of the reference CLI is out of scope.

    python -m graph_kmer_index_b200 make_from_flat -f flat.npz -o index [-m MODULO] [-S True] [-s True] [-M True] [-r True -k 31]
"""
import argparse
import logging
import sys

from .collision_free_kmer_index import CollisionFreeKmerIndex, MinimalKmerIndex
from .flat_kmers import FlatKmers
from .reference_kmer_index import ReferenceKmerIndex
from .reverse_kmer_index import ReverseKmerIndex


def make_from_flat(args):
    flat = FlatKmers.from_file(args.flat_index)
    if args.add_reverse_complements:
        logging.info("Will add reverse complements of every hash (k=%d)" % args.kmer_size)
        flat = FlatKmers.from_multiple_flat_kmers([flat, flat.get_reverse_complement_flat_kmers(k=args.kmer_size)])
    if args.make_minimal:
        index = MinimalKmerIndex.from_flat_kmers(flat, modulo=args.hash_modulo)
    else:
        index = CollisionFreeKmerIndex.from_flat_kmers(flat, modulo=args.hash_modulo, skip_frequencies=args.skip_frequencies,
                                                       skip_singletons=args.skip_singletons)
    index.to_file(args.out_file_name)
    logging.info("Done making kmer index")


def make_reverse(args):
    ReverseKmerIndex.from_flat_kmers(FlatKmers.from_file(args.flat_index)).to_file(args.out_file_name)
    logging.info("Done. Wrote reverse index to file: %s" % args.out_file_name)


def make_reference_kmer_index(args):
    if args.reference_fasta is not None:
        index = ReferenceKmerIndex.from_linear_reference(args.reference_fasta, args.reference_name, args.kmer_size, args.only_store_kmers)
    else:
        index = ReferenceKmerIndex.from_flat_kmers(FlatKmers.from_file(args.flat_index))
    index.to_file(args.out_file_name)
    logging.info("Saved reference kmer index to file %s" % args.out_file_name)


def add_reverse_complements(args):
    flat = FlatKmers.from_file(args.flat_kmers)
    flat = FlatKmers.from_multiple_flat_kmers([flat, flat.get_reverse_complement_flat_kmers(k=args.kmer_size)])
    flat.to_file(args.out_file_name)
    logging.info("Saved new flat kmers with reverse complements to %s" % args.out_file_name)


def run_argument_parser(args):
    parser = argparse.ArgumentParser(description="Graph Kmer Index (B200).", prog="graph_kmer_index_b200")
    subparsers = parser.add_subparsers()

    sub = subparsers.add_parser("make_from_flat")
    sub.add_argument("-o", "--out_file_name", required=True)
    sub.add_argument("-f", "--flat-index", required=True)
    sub.add_argument("-m", "--hash_modulo", required=False, type=int, default=452930477)
    sub.add_argument("-S", "--skip-frequencies", type=bool, default=False, required=False)
    sub.add_argument("-s", "--skip-singletons", type=bool, default=False, required=False)
    sub.add_argument("-M", "--make-minimal", type=bool, default=False, required=False)
    sub.add_argument("-r", "--add-reverse-complements", type=bool, default=False, required=False)
    sub.add_argument("-k", "--kmer-size", type=int, default=31, required=False)
    sub.set_defaults(func=make_from_flat)

    sub = subparsers.add_parser("make_reverse")
    sub.add_argument("-f", "--flat-index", required=True)
    sub.add_argument("-o", "--out-file-name", required=True)
    sub.set_defaults(func=make_reverse)

    sub = subparsers.add_parser("make_reference_kmer_index")
    sub.add_argument("-f", "--flat-index", required=False)
    sub.add_argument("-r", "--reference-fasta", required=False)
    sub.add_argument("-n", "--reference-name", required=False)
    sub.add_argument("-k", "--kmer-size", required=False, type=int, default=16)
    sub.add_argument("-o", "--out-file-name", required=True)
    sub.add_argument("-O", "--only-store-kmers", required=False, default=False, type=bool)
    sub.set_defaults(func=make_reference_kmer_index)

    sub = subparsers.add_parser("add_reverse_complements")
    sub.add_argument("-f", "--flat-kmers", required=True)
    sub.add_argument("-o", "--out-file-name", required=True)
    sub.add_argument("-k", "--kmer-size", type=int, required=True)
    sub.set_defaults(func=add_reverse_complements)

    if len(args) == 0:
        parser.print_help()
        sys.exit(1)
    parsed = parser.parse_args(args)
    parsed.func(parsed)


def main():
    logging.basicConfig(level=logging.INFO, format="%(asctime)s, %(levelname)s: %(message)s")
    run_argument_parser(sys.argv[1:])
