"""Host-side tokenizer wrapper: same surface as ``faster_whisper.tokenizer.Tokenizer``
(faster-whisper==1.2.0, a pip dependency of the reference; constructed at
whisper_live/transcriber/transcriber_faster_whisper.py:909 and
whisper_live/batch_inference.py:293, consumed at :981-1045, :1092, :1293, :1493,
:1532, :1671, :1837-1850).  Token ids are the only thing that crosses the C ABI;
text <-> ids stays in Python on top of HF ``tokenizers`` exactly as in the reference.

``build_synthetic_tokenizer`` exists because no tokenizer.json is available offline:
it fabricates a byte-level vocabulary with Whisper's special-token ids in the right
places so every code path (prompts, suppress list, word splitting, timestamps) runs.
"""
from __future__ import annotations

import string
from functools import cached_property
from typing import List, Optional, Tuple

LANGUAGE_CODES = (
    "en zh de es ru ko fr ja pt tr pl ca nl ar sv it id hi fi vi he uk el ms cs ro da hu ta no th ur hr bg lt la mi "
    "ml cy sk te fa lv bn sr az sl kn et mk br eu is hy ne mn bs kk sq sw gl mr pa si km sn yo so af oc ka be tg sd "
    "gu am yi lo uz fo ht ps tk nn mt sa lb my bo tl mg as tt haw ln ha ba jw su yue").split()
_LANGUAGE_CODES = tuple(LANGUAGE_CODES)
_TASKS = ("transcribe", "translate")
_NO_SPACE_LANGS = {"zh", "ja", "th", "lo", "my", "yue"}


class Tokenizer:
    """Whisper-specific view over a ``tokenizers.Tokenizer``."""

    def __init__(self, tokenizer, multilingual: bool, task: Optional[str] = None, language: Optional[str] = None):
        self.tokenizer = tokenizer
        if multilingual:
            if task not in _TASKS:
                raise ValueError(f"'{task}' is not a valid task (accepted tasks: {', '.join(_TASKS)})")
            if language not in _LANGUAGE_CODES:
                raise ValueError(f"'{language}' is not a valid language code (accepted language codes: "
                                 f"{', '.join(_LANGUAGE_CODES)})")
            self.task = self._id(f"<|{task}|>")
            self.language = self._id(f"<|{language}|>")
            self.language_code = language
        else:
            self.task = None
            self.language = None
            self.language_code = "en"

    def _id(self, token: str) -> int:
        i = self.tokenizer.token_to_id(token)
        if i is None:
            raise KeyError(f"token {token!r} is not in the vocabulary")
        return i

    @cached_property
    def transcribe(self) -> int:
        return self._id("<|transcribe|>")

    @cached_property
    def translate(self) -> int:
        return self._id("<|translate|>")

    @cached_property
    def sot(self) -> int:
        return self._id("<|startoftranscript|>")

    @cached_property
    def sot_lm(self) -> int:
        return self._id("<|startoflm|>")

    @cached_property
    def sot_prev(self) -> int:
        return self._id("<|startofprev|>")

    @cached_property
    def eot(self) -> int:
        return self._id("<|endoftext|>")

    @cached_property
    def no_timestamps(self) -> int:
        return self._id("<|notimestamps|>")

    @property
    def timestamp_begin(self) -> int:
        return self.no_timestamps + 1

    @property
    def sot_sequence(self) -> List[int]:
        seq = [self.sot]
        if self.language is not None:
            seq.append(self.language)
        if self.task is not None:
            seq.append(self.task)
        return seq

    def encode(self, text: str) -> List[int]:
        return self.tokenizer.encode(text, add_special_tokens=False).ids

    def decode(self, tokens: List[int]) -> str:
        return self.tokenizer.decode([t for t in tokens if t < self.eot])

    def decode_with_timestamps(self, tokens: List[int]) -> str:
        pieces: List[List[int]] = [[]]
        for t in tokens:
            if t >= self.timestamp_begin:
                pieces.append(f"<|{(t - self.timestamp_begin) * 0.02:.2f}|>")
                pieces.append([])
            else:
                pieces[-1].append(t)
        return "".join(p if isinstance(p, str) else self.tokenizer.decode(p) for p in pieces)

    @cached_property
    def non_speech_tokens(self) -> Tuple[int, ...]:
        """Ids of symbols that annotate non-speech (brackets, music notes, ...); the
        default ``suppress_tokens=[-1]`` set (transcriber_faster_whisper.py:1835-1837)."""
        symbols = list('"#()*+/:;<=>@[\\]^_`{|}~「」『』')
        symbols += "<< >> <<< >>> -- --- -( -[ (' (\" (( )) ((( ))) [[ ]] {{ }} ♪♪ ♪♪♪".split()
        # these need to be handled separately: they may be split into several byte tokens
        misc = set("♩♪♫♬♭♮♯")
        out = {self.encode(" -")[0], self.encode(" '")[0]}
        for sym in symbols + list(misc):
            for ids in (self.encode(sym), self.encode(" " + sym)):
                if len(ids) == 1 or sym in misc:
                    out.add(ids[0])
        return tuple(sorted(out))

    def split_to_word_tokens(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        if self.language_code in _NO_SPACE_LANGS:
            return self.split_tokens_on_unicode(tokens)
        return self.split_tokens_on_spaces(tokens)

    def split_tokens_on_unicode(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        full = self.decode_with_timestamps(tokens)
        bad = "�"
        words, word_tokens, cur, offset = [], [], [], 0
        for t in tokens:
            cur.append(t)
            text = self.decode_with_timestamps(cur)
            try:
                pos = offset + text.index(bad)
                complete = pos < len(full) and full[pos] == bad
            except ValueError:
                complete = True
            if complete:
                words.append(text)
                word_tokens.append(cur)
                cur = []
                offset += len(text)
        return words, word_tokens

    def split_tokens_on_spaces(self, tokens: List[int]) -> Tuple[List[str], List[List[int]]]:
        subwords, subword_tokens = self.split_tokens_on_unicode(tokens)
        words: List[str] = []
        word_tokens: List[List[int]] = []
        for sw, st in zip(subwords, subword_tokens):
            starts_word = (st[0] >= self.eot) or sw.startswith(" ") or (sw.strip() in string.punctuation) or not words
            if starts_word:
                words.append(sw)
                word_tokens.append(st)
            else:
                words[-1] += sw
                word_tokens[-1].extend(st)
        return words, word_tokens


def _gpt2_byte_chars() -> List[str]:
    """The byte-level BPE alphabet in GPT-2 id order (id 220 is 'Ġ', the space)."""
    printable = list(range(33, 127)) + list(range(161, 173)) + list(range(174, 256))
    chars = [chr(b) for b in printable]
    extra = 0
    for b in range(256):
        if b not in printable:
            chars.append(chr(256 + extra))
            extra += 1
    return chars


def build_synthetic_tokenizer(vocab_size: int):
    """A ``tokenizers.Tokenizer`` with Whisper's id layout and a fabricated text vocabulary:
    ids 0..255 are the byte alphabet, 256..eot-1 decode to ' w<id>', specials sit where the
    released vocabularies put them."""
    import tokenizers
    from tokenizers import decoders, models, pre_tokenizers

    multilingual = vocab_size >= 51865
    eot = 50257 if multilingual else 50256
    n_lang = vocab_size - 51765 - 1 if multilingual else 99
    vocab = {c: i for i, c in enumerate(_gpt2_byte_chars())}
    for i in range(256, eot):
        vocab[f"Ġw{i}"] = i
    tok = tokenizers.Tokenizer(models.BPE(vocab=vocab, merges=[]))
    tok.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)
    tok.decoder = decoders.ByteLevel()
    specials = ["<|endoftext|>", "<|startoftranscript|>"]
    specials += [f"<|{c}|>" for c in LANGUAGE_CODES[:n_lang]]
    specials += ["<|translate|>", "<|transcribe|>", "<|startoflm|>", "<|startofprev|>",
                 "<|nospeech|>" if multilingual else "<|nocaptions|>", "<|notimestamps|>"]
    specials += [f"<|{i * 0.02:.2f}|>" for i in range(1501)]
    tok.add_special_tokens([tokenizers.AddedToken(s, special=True) for s in specials])
    assert tok.token_to_id("<|endoftext|>") == eot, tok.token_to_id("<|endoftext|>")
    assert tok.get_vocab_size() == vocab_size, (tok.get_vocab_size(), vocab_size)
    return tok
