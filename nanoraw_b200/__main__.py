"""`python -m nanoraw_b200 genome_resquiggle ...` -- the reference's `nanoraw` console script
(__main__.py:13-112, setup.py:21-25) restricted to the sub-command this package replaces."""
import argparse
import sys

from . import option_parsers


def main(args=None):
    commands = [('genome_resquiggle',
                 'Re-annotate raw signal to match genomic alignment and save corrected events.',
                 option_parsers.get_resquiggle_parser())]
    parser = argparse.ArgumentParser(
        prog='nanoraw',
        description='nanoraw (B200 build): only genome_resquiggle is provided; plotting, '
                    'statistics and wiggle export are served by the reference package.')
    sub = parser.add_subparsers(title='commands', dest='command')
    sub.required = True
    for name, text, cmd_parser in commands:
        sub.add_parser(name, parents=[cmd_parser], help=text, add_help=False,
                       description=cmd_parser.description)
    ns = parser.parse_args(args)
    from .pipeline import resquiggle_main
    resquiggle_main(ns)


if __name__ == '__main__':
    main(sys.argv[1:])
