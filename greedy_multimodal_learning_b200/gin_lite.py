"""Minimal gin-config compatible front end.

`gin` and `argh` are not installed in this image and there is no network.  The reference's five
config files only use the plainest gin syntax -- `Name.param = <python literal>` lines and `#`
comments (configs/*.gin) -- and its CLI is `train.py <save_path> <cfg>[#cfg2] [bindings]`
(src/utils.py:58-68).  This module implements exactly that subset:

    @configurable                      # or @configurable("OtherName")
    def train(save_path, lr, ...): ...
    parse_config_files_and_bindings(["configs/training_guided.gin"], "train.lr=0.01")
    train("out/")                      # unbound arguments are filled from the parsed bindings

Unsupported gin features (macros `%x`, references `@x`, scopes `a/b.c`, imports) raise
NotImplementedError instead of being silently ignored.
"""
from __future__ import annotations

import ast
import functools
import inspect
from typing import Any, Dict, Iterable, Tuple

_BINDINGS: Dict[Tuple[str, str], Any] = {}
_REGISTRY: Dict[str, Any] = {}


def clear_config():
    _BINDINGS.clear()


def bind_parameter(key: str, value):
    name, param = key.rsplit(".", 1)
    _BINDINGS[(name.strip(), param.strip())] = value


def query_parameter(key: str):
    name, param = key.rsplit(".", 1)
    return _BINDINGS[(name, param)]


def config_dict() -> Dict[str, Any]:
    """Flat `{'Name.param': value}` view (what the reference passes around as `_CONFIG`)."""
    return {"%s.%s" % k: v for k, v in _BINDINGS.items()}


def _strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            out.append(ch)
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
            out.append(ch)
        elif ch == "#":
            break
        else:
            out.append(ch)
    return "".join(out).strip()


def parse_config(text: str):
    """Parse `Name.param = literal` statements; a value may span lines while brackets are open."""
    pending = ""
    for raw in text.splitlines():
        line = _strip_comment(raw)
        if not line and not pending:
            continue
        pending = (pending + " " + line).strip() if pending else line
        if pending.count("[") + pending.count("(") + pending.count("{") > \
                pending.count("]") + pending.count(")") + pending.count("}"):
            continue
        stmt, pending = pending, ""
        if stmt.startswith(("import ", "include ")):
            raise NotImplementedError("gin_lite: '%s' is not supported" % stmt.split()[0])
        if "=" not in stmt:
            raise ValueError("gin_lite: cannot parse %r" % stmt)
        key, value = stmt.split("=", 1)
        key, value = key.strip(), value.strip()
        if "/" in key or value.startswith(("@", "%")):
            raise NotImplementedError("gin_lite: scopes, references and macros are not supported (%r)" % stmt)
        if "." not in key:
            raise ValueError("gin_lite: binding key must be Name.param (%r)" % key)
        bind_parameter(key, ast.literal_eval(value))
    if pending:
        raise ValueError("gin_lite: unterminated value %r" % pending)


def parse_config_files_and_bindings(config_files: Iterable[str], bindings=""):
    """Same call as the reference's `gin.parse_config_files_and_bindings` (src/utils.py:61); bindings may be
    a newline- or ';'-separated string or a list of strings and override the files."""
    for path in config_files or []:
        if path:
            with open(path) as f:
                parse_config(f.read())
    if not isinstance(bindings, str):
        bindings = "\n".join(bindings or [])
    parse_config(bindings.replace(";", "\n"))


def configurable(obj=None):
    """Decorator: arguments the caller leaves out are taken from the parsed bindings."""

    def wrap(target, name=None):
        name = name or target.__name__
        if inspect.isclass(target):
            orig = target.__init__

            @functools.wraps(orig)
            def __init__(self, *a, **k):
                orig(self, *a, **_fill(name, orig, a, k, skip_self=True))

            target.__init__ = __init__
            _REGISTRY[name] = target
            return target

        @functools.wraps(target)
        def fn(*a, **k):
            return target(*a, **_fill(name, target, a, k))

        _REGISTRY[name] = fn
        return fn

    if isinstance(obj, str):
        return lambda t: wrap(t, obj)
    if obj is not None:
        return wrap(obj)
    return wrap


def external_configurable(target, name=None):
    return configurable(name)(target) if name else configurable(target)


def _fill(name, fn, args, kwargs, skip_self=False):
    params = list(inspect.signature(fn).parameters)
    if skip_self:
        params = params[1:]
    given = set(params[:len(args)]) | set(kwargs)
    out = dict(kwargs)
    for (n, p), v in _BINDINGS.items():
        if n == name and p not in given:
            if p not in params and not any(
                    q.kind == inspect.Parameter.VAR_KEYWORD for q in inspect.signature(fn).parameters.values()):
                raise TypeError("gin_lite: %s has no parameter %r" % (name, p))
            out[p] = v
    return out
