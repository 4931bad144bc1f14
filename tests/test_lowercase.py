"""Lower-casing (src/search/search_field.rs:284,312: Rust's `str::to_lowercase` on the query and on every matched term
before scoring).  Three implementations: CPython's `str.lower()` (the same Unicode algorithm -- full mappings plus the
final-sigma rule -- written by someone else: the pin), the oracle's (oracle/rust_lower.hpp, tables derived from the general
categories) and the product's (csrc/format/unicode.hpp, tables derived by probing).  Over every scalar in three contexts,
the sigma contexts, and random strings."""
import ctypes
import random

import pytest

import helpers


@pytest.fixture(scope="module")
def lowerers(native_libs):
    oracle = helpers.Oracle()
    lib = helpers._index_lib()
    lib.vidx_to_lowercase.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]

    def product(text):
        out = ctypes.create_string_buffer(len(text.encode("utf-8")) * 3 + 16)
        assert lib.vidx_to_lowercase(text.encode("utf-8"), out, len(out)) == 0
        return out.value.decode("utf-8")

    return (lambda text: oracle.call("to_lowercase", text=text)), product


SCALARS = [cp for cp in range(1, 0x110000) if not 0xD800 <= cp <= 0xDFFF]


def test_vectors(lowerers):
    for lower in lowerers:
        assert lower("İstanbul") == "i̇stanbul" and len(lower("İ")) == 2          # U+0130 expands
        assert lower("ΟΔΟΣ ΑΘΗΝΑΣ") == "οδος αθηνας"                              # word-final sigma
        assert lower("Σ") == "σ" and lower("ΑΣΑ") == "ασα" and lower("Α.Σ") == "α.ς" and lower("ΑΣ.Α") == "ας.α".replace("ς", "σ")
        assert lower("ΑΣ'") == "ας'" and lower("'Σ") == "'σ"                       # case-ignorable scalars are skipped
        assert lower("Straße STRASSE ẞ") == "straße strasse ß"
        assert lower("KÅ") == "kå"                                               # Kelvin and Angstrom signs
        assert lower("ǅ") == "ǆ" and lower("食べる Ａ") == "食べる ａ"


def test_every_scalar_in_context(lowerers):
    oracle_lower, product_lower = lowerers
    # in bulk: strings of 256 scalars separated by a neutral scalar, alone and after / before a cased letter and a sigma
    for make in (lambda c: c + "|", lambda c: "a" + c + "b|", lambda c: "aΣ" + c + "|", lambda c: c + "Σ|", lambda c: "aΣ" + c + "a|"):
        for base in range(0, len(SCALARS), 4096):
            text = "".join(make(chr(cp)) for cp in SCALARS[base:base + 4096])
            want = text.lower()
            assert oracle_lower(text) == want, (hex(SCALARS[base]),)
            assert product_lower(text) == want, (hex(SCALARS[base]),)


def test_random_strings(lowerers):
    oracle_lower, product_lower = lowerers
    rng = random.Random(3)
    pool = list("abcXYZ Σσς.'’:-İIıiẞßÅKǅ食̇­ʰ") + [chr(cp) for cp in rng.sample(SCALARS, 300)]
    for _ in range(3000):
        text = "".join(rng.choice(pool) for _ in range(rng.randint(0, 12)))
        want = text.lower()
        assert oracle_lower(text) == want, [hex(ord(c)) for c in text]
        assert product_lower(text) == want, [hex(ord(c)) for c in text]
