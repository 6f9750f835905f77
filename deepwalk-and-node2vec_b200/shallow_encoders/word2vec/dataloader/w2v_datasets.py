"""Registered text datasets (reference: shallow_encoders/word2vec/dataloader/w2v_datasets.py:13-102): the two in-memory toy corpora
(`test`, `abcde` -- their sentences are the contract: `abcde` must teach a~b, c~d, e alone) and the file-backed corpora, which need
assets that are not shipped (`assets/wikitext-*/wiki.train.tokens`, `assets/Shakespeare_data.csv`; tools/download_dataset.sh of
the reference fetches them) and raise FileNotFoundError when iterated without them."""
import os

from shallow_encoders.common.path import ASSETS_PATH
from shallow_encoders.word2vec.dataloader.iterators import FileIterator, InMemoryIterator
from shallow_encoders.word2vec.dataloader.registry import register_dataset


@register_dataset('test')
class TestDataset(InMemoryIterator):
    """Four tiny sentences (punctuation, repeats, an empty one) that exercise the tokenizer and the vocabulary."""
    __test__ = False                      # not a pytest class

    def __init__(self):
        super().__init__(['a, a, c, b, b', 'hello world! hello world!', 'test here, test there, here there', '.'])


@register_dataset('abcde')
class ABCDEDataset(InMemoryIterator):
    """`a` goes with `b`, `c` with `d`, `e` alone."""

    def __init__(self):
        ab = ['a b a b a b a b a b', 'a b a b a b', 'b a b a', 'a b a b a b a b']
        cd = ['c d c d c d c d', 'd c d c d c', 'c d c d c d']
        e = ['e e e e e e e e', 'e e e']
        super().__init__(ab + cd + e)


class WikiTextDataset(FileIterator):
    """`<assets>/<dataset_name>/wiki.<split>.tokens`, one paragraph per line."""

    def __init__(self, dataset_name: str, split: str, assets_path: str = ASSETS_PATH):
        super().__init__(os.path.join(assets_path, dataset_name, f'wiki.{split}.tokens'))


@register_dataset('wiki-text-2')
class WikiText2Dataset(WikiTextDataset):
    def __init__(self, *args, **kwargs):
        super().__init__('wikitext-2', 'train', *args, **kwargs)


@register_dataset('wiki-text-103')
class WikiText103Dataset(WikiTextDataset):
    def __init__(self, *args, **kwargs):
        super().__init__('wikitext-103', 'train', *args, **kwargs)


@register_dataset('shakespeare')
class ShakespeareDataset(InMemoryIterator):
    """The `PlayerLine` column of `<assets>/Shakespeare_data.csv`."""

    def __init__(self, assets_path: str = ASSETS_PATH):
        import pandas as pd
        super().__init__(pd.read_csv(os.path.join(assets_path, 'Shakespeare_data.csv'))['PlayerLine'].values.tolist())
