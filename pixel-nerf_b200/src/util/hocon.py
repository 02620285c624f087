"""
HOCON-subset reader standing in for `pyhocon.ConfigFactory` (not installed here; the
reference loads conf/*.conf through it: src/util/args.py:6,89-99).

Supports exactly what the reference's conf files use: `#`/`//` comments, `key = value` and
`key : value`, nested `key { ... }` blocks (re-opening a block merges into it), dotted keys,
`[a, b, ...]` lists (possibly nested / multi-line), quoted and bare strings, ints, floats,
true/false (any case), null, and `include required("relative/path")` / `include "path"`
resolved relative to the including file.  Later definitions override earlier ones.

ConfigTree mirrors the accessors the reference calls:
`conf[key]` (dotted ok), `key in conf`, `get(key, default)`, and
`get_int/get_float/get_bool/get_string/get_list(key[, default])`.
"""
import os
import re

_UNSET = object()


class ConfigMissingException(KeyError):
    pass


class ConfigTree(dict):
    def _lookup(self, key):
        cur = self
        for part in str(key).split("."):
            if not isinstance(cur, dict) or not dict.__contains__(cur, part):
                return _UNSET
            cur = dict.__getitem__(cur, part)
        return cur

    def __getitem__(self, key):
        v = self._lookup(key)
        if v is _UNSET:
            raise ConfigMissingException(f"No configuration setting found for key {key}")
        return v

    def __contains__(self, key):
        return self._lookup(key) is not _UNSET

    def get(self, key, default=_UNSET):
        v = self._lookup(key)
        if v is _UNSET:
            if default is _UNSET:
                raise ConfigMissingException(f"No configuration setting found for key {key}")
            return default
        return v

    def _typed(self, key, default, conv):
        v = self._lookup(key)
        if v is _UNSET:
            if default is _UNSET:
                raise ConfigMissingException(f"No configuration setting found for key {key}")
            return default
        return None if v is None else conv(v)

    def get_int(self, key, default=_UNSET):
        return self._typed(key, default, int)

    def get_float(self, key, default=_UNSET):
        return self._typed(key, default, float)

    def get_string(self, key, default=_UNSET):
        def conv(v):
            if isinstance(v, bool):
                return "true" if v else "false"
            return str(v)
        return self._typed(key, default, conv)

    def get_bool(self, key, default=_UNSET):
        def conv(v):
            if isinstance(v, bool):
                return v
            s = str(v).strip().lower()
            if s in ("true", "yes", "on", "1"):
                return True
            if s in ("false", "no", "off", "0"):
                return False
            raise ValueError(f"{key}: '{v}' is not a boolean")
        return self._typed(key, default, conv)

    def get_list(self, key, default=_UNSET):
        def conv(v):
            if not isinstance(v, list):
                raise ValueError(f"{key} is not a list")
            return v
        return self._typed(key, default, conv)

    def get_config(self, key, default=_UNSET):
        return self._typed(key, default, lambda v: v)

    def put(self, key, value):
        parts = str(key).split(".")
        cur = self
        for p in parts[:-1]:
            nxt = dict.get(cur, p)
            if not isinstance(nxt, ConfigTree):
                nxt = ConfigTree()
                dict.__setitem__(cur, p, nxt)
            cur = nxt
        old = dict.get(cur, parts[-1])
        if isinstance(old, ConfigTree) and isinstance(value, ConfigTree):
            _merge(old, value)
        else:
            dict.__setitem__(cur, parts[-1], value)


def _merge(dst, src):
    for k, v in src.items():
        old = dict.get(dst, k)
        if isinstance(old, ConfigTree) and isinstance(v, ConfigTree):
            _merge(old, v)
        else:
            dict.__setitem__(dst, k, v)


_TOKEN = re.compile(r"""
    (?P<ws>[ \t\r]+) | (?P<nl>\n) | (?P<comment>(\#|//)[^\n]*) |
    (?P<str>"(?:[^"\\]|\\.)*") | (?P<punct>[{}\[\],=:]) |
    (?P<bare>[^\s{}\[\],=:"\#]+)
""", re.X)


def _tokenize(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f"HOCON: cannot tokenize at offset {pos}: {text[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind in ("ws", "comment"):
            continue
        out.append((kind, m.group(kind)))
    out.append(("eof", ""))
    return out


def _scalar(tok):
    low = tok.lower()
    if low == "true":
        return True
    if low == "false":
        return False
    if low == "null":
        return None
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


class _Parser:
    def __init__(self, text, basedir):
        self.t = _tokenize(text)
        self.i = 0
        self.basedir = basedir

    def peek(self):
        return self.t[self.i]

    def next(self):
        tok = self.t[self.i]
        self.i += 1
        return tok

    def skip_sep(self):
        while self.peek()[0] == "nl" or self.peek() == ("punct", ","):
            self.i += 1

    def parse_object(self, until_brace):
        tree = ConfigTree()
        while True:
            self.skip_sep()
            kind, val = self.peek()
            if kind == "eof":
                if until_brace:
                    raise ValueError("HOCON: missing '}'")
                return tree
            if (kind, val) == ("punct", "}"):
                if not until_brace:
                    raise ValueError("HOCON: unexpected '}'")
                self.next()
                return tree
            if kind == "bare" and val == "include":
                self.next()
                self.parse_include(tree)
                continue
            if kind not in ("bare", "str"):
                raise ValueError(f"HOCON: expected a key, got {val!r}")
            self.next()
            key = val[1:-1] if kind == "str" else val
            kind2, val2 = self.peek()
            if (kind2, val2) == ("punct", "{"):
                self.next()
                tree.put(key, self.parse_object(True))
                continue
            if kind2 == "punct" and val2 in "=:":
                self.next()
                while self.peek()[0] == "nl":
                    self.next()
                tree.put(key, self.parse_value())
                continue
            raise ValueError(f"HOCON: expected '=', ':' or '{{' after key {key!r}")

    def parse_include(self, tree):
        kind, val = self.next()
        path = None
        if kind == "bare" and val.startswith("required("):
            # tokenizer splits required("x") as bare 'required(' + str + bare ')'
            kind, sval = self.next()
            path = sval[1:-1]
            self.next()  # ')'
        elif kind == "bare" and val in ("required", "file", "url"):
            raise ValueError("HOCON: unsupported include form")
        elif kind == "str":
            path = val[1:-1]
        if path is None:
            raise ValueError("HOCON: malformed include")
        full = path if os.path.isabs(path) else os.path.join(self.basedir, path)
        _merge(tree, parse_file(full))

    def parse_value(self):
        kind, val = self.next()
        if (kind, val) == ("punct", "{"):
            return self.parse_object(True)
        if (kind, val) == ("punct", "["):
            items = []
            while True:
                self.skip_sep()
                if self.peek() == ("punct", "]"):
                    self.next()
                    return items
                items.append(self.parse_value())
        if kind == "str":
            return bytes(val[1:-1], "utf-8").decode("unicode_escape")
        if kind == "bare":
            # bare values may continue with more bare words on the same line ("a b c")
            parts = [val]
            while self.peek()[0] == "bare":
                parts.append(self.next()[1])
            return _scalar(parts[0]) if len(parts) == 1 else " ".join(parts)
        raise ValueError(f"HOCON: unexpected token {val!r} in value position")


def parse_string(text, basedir="."):
    return _Parser(text, basedir).parse_object(False)


def parse_file(path):
    with open(path, "r") as f:
        return parse_string(f.read(), os.path.dirname(os.path.abspath(path)))


def from_dict(d):
    t = ConfigTree()
    for k, v in d.items():
        t.put(k, from_dict(v) if isinstance(v, dict) else v)
    return t


class ConfigFactory:
    """pyhocon.ConfigFactory look-alike."""
    parse_file = staticmethod(parse_file)
    parse_string = staticmethod(parse_string)
    from_dict = staticmethod(from_dict)
