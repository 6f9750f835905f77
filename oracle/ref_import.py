"""
TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Import shim for the UNMODIFIED reference: /root/reference (read-only, present only in the build
container) or its byte-for-byte, git-ignored copy under baseline/_ref/ (oracle/vendor_ref.py), which
travels to the GPU box.  Used by `oracle/make_golden.py` to (a) validate the oracle restatements against
the real reference and (b) generate the golden vectors committed under `tests/golden/`, and by
`oracle/ref_pipeline.py` (bench.py's `--impl reference` / `cpu_baseline` legs) to TIME the real reference.

Three third-party packages the reference imports are not installable offline, so in-memory
stubs with the documented behaviour are registered before the reference modules load:

  * nltk.stem.WordNetLemmatizer        (shallow_encoders/word2vec/dataloader/torch_dataset.py:12)
    identity lemmatizer -- graph datasets never lemmatise (torch_dataset.py:230).
  * torchtext.vocab.build_vocab_from_iterator   (torch_dataset.py:14, :104-110)
    torchtext 0.15.2 ordering: specials first, then tokens by (-frequency, token).
  * pytorch_lightning.LightningModule  (shallow_encoders/word2vec/trainer.py:5, :18)
    nn.Module with a no-op `log`.
"""
import os
import sys
import types
from collections import Counter

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VENDORED_ROOT = os.path.join(_REPO, 'baseline', '_ref')      # git-ignored byte-for-byte copy (oracle/vendor_ref.py)


def reference_root():
    """/root/reference where it is mounted (the build container), else the vendored copy that travels to the GPU box."""
    for cand in (os.environ.get('SE_REFERENCE_ROOT'), '/root/reference', VENDORED_ROOT):
        if cand and os.path.isdir(os.path.join(cand, 'shallow_encoders')):
            return cand
    return None


REFERENCE_ROOT = reference_root() or '/root/reference'


class _StubVocab:
    def __init__(self, itos):
        self._itos = list(itos)
        self._stoi = {t: i for i, t in enumerate(self._itos)}
        self._default = None

    def __len__(self):
        return len(self._itos)

    def __contains__(self, token):
        return token in self._stoi

    def __getitem__(self, token):
        if token in self._stoi:
            return self._stoi[token]
        if self._default is None:
            raise RuntimeError(f'Token {token} not found and default index is not set')
        return self._default

    def __call__(self, tokens):
        return [self[t] for t in tokens]

    def set_default_index(self, index):
        self._default = index

    def get_stoi(self):
        return dict(self._stoi)

    def get_itos(self):
        return list(self._itos)


def _build_vocab_from_iterator(iterator, min_freq=1, specials=None, special_first=True, max_tokens=None):
    counter = Counter()
    for tokens in iterator:
        counter.update(tokens)
    specials = list(specials or [])
    for s in specials:
        counter.pop(s, None)
    ordered = sorted(counter.items(), key=lambda kv: (-kv[1], kv[0]))
    tokens = [t for t, f in ordered if f >= min_freq]
    if max_tokens is not None:
        tokens = tokens[:max_tokens - len(specials)]
    itos = specials + tokens if special_first else tokens + specials
    return _StubVocab(itos)


def install_stubs():
    import torch

    if 'nltk' not in sys.modules:
        nltk = types.ModuleType('nltk')
        stem = types.ModuleType('nltk.stem')

        class WordNetLemmatizer:
            def lemmatize(self, word, pos='n'):
                return word

        stem.WordNetLemmatizer = WordNetLemmatizer
        nltk.stem = stem
        sys.modules['nltk'] = nltk
        sys.modules['nltk.stem'] = stem

    if 'torchtext' not in sys.modules:
        torchtext = types.ModuleType('torchtext')
        vocab = types.ModuleType('torchtext.vocab')
        vocab.build_vocab_from_iterator = _build_vocab_from_iterator
        torchtext.vocab = vocab
        sys.modules['torchtext'] = torchtext
        sys.modules['torchtext.vocab'] = vocab

    if 'pytorch_lightning' not in sys.modules:
        pl = types.ModuleType('pytorch_lightning')

        class LightningModule(torch.nn.Module):
            def log(self, *args, **kwargs):
                pass

        pl.LightningModule = LightningModule
        sys.modules['pytorch_lightning'] = pl


def import_reference():
    """Put the reference on sys.path (ahead of this repo's own `shallow_encoders` mirror)
    and return its `shallow_encoders` package.  Raises if the reference is not mounted."""
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'shallow_encoders')):
        raise RuntimeError(f'reference not mounted at {REFERENCE_ROOT}')
    sys.dont_write_bytecode = True
    install_stubs()
    # make sure a previously imported mirror package does not shadow the reference
    for name in [m for m in sys.modules if m == 'shallow_encoders' or m.startswith('shallow_encoders.')]:
        del sys.modules[name]
    sys.path.insert(0, REFERENCE_ROOT)
    import shallow_encoders  # noqa: F401
    return shallow_encoders
