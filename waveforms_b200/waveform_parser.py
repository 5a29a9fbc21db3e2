"""``wave_eval``: string -> Waveform.

Drop-in for /root/reference/waveforms/waveform_parser.py:255-315.  The
reference parses with an ANTLR4-generated lexer/parser (generated files are not
in its repository and need Java); this is a hand-written lexer + precedence
climbing parser that follows the grammar /root/reference/waveforms/Waveform.g4
rule by rule, including what ANTLR4's left-recursion rewrite does with it:

* binary operators, highest precedence first (Waveform.g4:9-12):
  ``** ^``  >  ``* /``  >  ``+ -``  >  ``<< >>``; all LEFT associative
  (ANTLR's default, so ``2**3**2 == 64``);
* unary minus is the 6th alternative (Waveform.g4:14), i.e. it binds LOOSER
  than every binary operator: its operand is parsed at precedence 8 and swallows
  all following binary operators (``-a + b`` is ``-(a + b)``, ``-1/2`` is
  ``-(1/2)``);
* lexer: longest match, ties to the earlier rule, so ``e``/``pi``/``inf`` are
  CONSTANT tokens and ``exp`` is an ID (Waveform.g4:43-57).

The evaluation side mirrors the reference visitor (waveform_parser.py:26-211):
names resolve with ``getattr`` on the ``waveform`` then ``multy_drag`` modules,
numbers and strings go through ``ast.literal_eval``, a numeric result is wrapped
in ``const`` and the result is ``simplify()``-ed; every failure surfaces as
``SyntaxError``.
"""
from __future__ import annotations

import re
from ast import literal_eval
from functools import lru_cache

from . import multy_drag, waveform


class WaveformParseError(Exception):
    """Custom exception for waveform parsing errors."""


_REAL = r'(?:\d+(?:\.\d*)?|\.\d+)(?:[eE][+-]?\d+)?'
_TOKEN = re.compile(
    r'\s*(?:'
    rf'(?P<NUMBER>{_REAL}j?)'
    r'|(?P<STRING>"[^"\r\n]*"|\'[^\'\r\n]*\')'
    r'|(?P<ID>[a-zA-Z_][a-zA-Z0-9_]*)'
    r'|(?P<OP>\*\*|<<|>>|[-+*/^()\[\],=])'
    r')')
_CONSTANTS = {'pi': waveform.pi, 'e': waveform.e, 'inf': waveform.inf}

# precedence numbers as ANTLR assigns them for the 13 alternatives of
# `expression` (alternative k gets 14 - k)
_BINARY = {'**': 13, '^': 13, '*': 12, '/': 12, '+': 11, '-': 11, '<<': 10,
           '>>': 10}
_UNARY_MINUS_OPERAND = 8


def _tokenize(text):
    tokens = []
    pos = 0
    end = len(text.rstrip())
    while pos < end:
        m = _TOKEN.match(text, pos)
        if m is None or m.end() == pos:
            raise WaveformParseError(
                f"Syntax error at line 1, column {pos}: token recognition "
                f"error at: '{text[pos:pos + 1]}'")
        kind = m.lastgroup
        value, col = m.group(kind), m.start(kind)
        if kind == 'ID' and value in _CONSTANTS:
            kind = 'CONSTANT'
        tokens.append((kind, value, col))
        pos = m.end()
    tokens.append(('EOF', '<EOF>', end))
    return tokens


class _Parser:

    def __init__(self, text):
        self.toks = _tokenize(text)
        self.i = 0

    # -- token helpers -----------------------------------------------------
    def peek(self, k=0):
        return self.toks[min(self.i + k, len(self.toks) - 1)]

    def at_op(self, *ops, k=0):
        kind, value, _ = self.peek(k)
        return kind == 'OP' and value in ops

    def advance(self):
        tok = self.toks[self.i]
        self.i += 1
        return tok

    def expect_op(self, op):
        if not self.at_op(op):
            self.fail(f"expecting '{op}'")
        return self.advance()

    def fail(self, what):
        kind, value, col = self.peek()
        raise WaveformParseError(
            f"Syntax error at line 1, column {col}: {what} at '{value}'")

    # -- grammar ----------------------------------------------------------------
    def parse(self):
        # expr: assignment | expression
        if self.peek()[0] == 'ID' and self.at_op('=', k=1):
            raise WaveformParseError(
                "Assignment expressions are not supported")
        value = self.expression(0)
        if self.peek()[0] != 'EOF':
            self.fail('extraneous input')
        return value

    def expression(self, min_prec):
        left = self.prefix()
        while True:
            kind, op, _ = self.peek()
            prec = _BINARY.get(op) if kind == 'OP' else None
            if prec is None or prec < min_prec:
                return left
            self.advance()
            right = self.expression(prec + 1)  # left associative
            left = self.binary(op, left, right)

    @staticmethod
    def binary(op, a, b):
        if op in ('**', '^'):
            return a**b
        if op == '*':
            return a * b
        if op == '/':
            return a / b
        if op == '+':
            return a + b
        if op == '-':
            return a - b
        if op == '<<':
            return a << b
        return a >> b

    def prefix(self):
        kind, value, _ = self.peek()
        if kind == 'OP' and value == '(':
            return self.parens_or_tuple()
        if kind == 'OP' and value == '-':
            self.advance()
            return -self.expression(_UNARY_MINUS_OPERAND)
        if kind == 'OP' and value == '[':
            return self.list_literal()
        if kind == 'CONSTANT':
            self.advance()
            return _CONSTANTS[value]
        if kind in ('NUMBER', 'STRING'):
            self.advance()
            return literal_eval(value)
        if kind == 'ID':
            if self.at_op('(', k=1):
                return self.function_call()
            raise WaveformParseError(f"Unknown identifier '{value}'")
        self.fail('no viable alternative')

    def parens_or_tuple(self):
        self.expect_op('(')
        first = self.expression(0)
        if self.at_op(')'):
            self.advance()
            return first
        items = [first]
        self.expect_op(',')
        while not self.at_op(')'):
            items.append(self.expression(0))
            if self.at_op(','):
                self.advance()
                if self.at_op(')') and len(items) > 1:
                    self.fail('trailing comma')  # grammar allows it only for 1-tuples
            elif not self.at_op(')'):
                self.fail("expecting ',' or ')'")
        self.advance()
        return tuple(items)

    def list_literal(self):
        self.expect_op('[')
        items = []
        if not self.at_op(']'):
            items.append(self.expression(0))
            while self.at_op(','):
                self.advance()
                items.append(self.expression(0))
        self.expect_op(']')
        return items

    def function_call(self):
        _, name, _ = self.advance()
        func = _resolve(name)
        self.expect_op('(')
        args, kwargs = [], {}
        while not self.at_op(')'):
            if self.peek()[0] == 'ID' and self.at_op('=', k=1):
                key = self.advance()[1]
                self.advance()
                kwargs[key] = self.expression(0)
            else:
                if kwargs:
                    self.fail('positional argument after keyword argument')
                args.append(self.expression(0))
            if self.at_op(','):
                self.advance()
                if self.at_op(')'):
                    self.fail('trailing comma')
            elif not self.at_op(')'):
                self.fail("expecting ',' or ')'")
        self.advance()
        return func(*args, **kwargs)


def _resolve(name):
    """waveform_parser.py:43-50 — any attribute of the two modules resolves."""
    for mod in (waveform, multy_drag):
        try:
            return getattr(mod, name)
        except AttributeError:
            continue
    raise WaveformParseError(f"Unknown function '{name}'")


def parse_waveform_expression(expr: str) -> waveform.Waveform:
    try:
        result = _Parser(expr).parse()
        if isinstance(result, (int, float, complex)):
            result = waveform.const(result)
        return result.simplify()
    except WaveformParseError:
        raise
    except Exception as exc:
        raise WaveformParseError(
            f"Failed to parse expression '{expr}': {str(exc)}")


@lru_cache(maxsize=1024)
def wave_eval(expr: str) -> "waveform.Waveform":
    """Parse and evaluate a waveform expression; raises SyntaxError on any
    failure (reference waveform_parser.py:296-315)."""
    try:
        return parse_waveform_expression(expr)
    except Exception as exc:
        raise SyntaxError(f"Failed to parse expression '{expr}': {str(exc)}")
