"""Sentence sources for text corpora (reference: shallow_encoders/word2vec/dataloader/iterators.py:7-58): restartable iterables of
raw sentence strings, in memory or line by line from a file."""
from typing import Iterator, List


class InMemoryIterator:
    """Iterates over a list of sentences; every `iter()` starts over."""

    def __init__(self, sentences: List[str]):
        self._sentences = list(sentences)
        self._cursor: Iterator[str] = iter(())

    def __iter__(self) -> 'InMemoryIterator':
        self._cursor = iter(self._sentences)
        return self

    def __next__(self) -> str:
        return next(self._cursor)

    def __len__(self) -> int:
        return len(self._sentences)


class FileIterator:
    """One sentence per line of a UTF-8 text file, read lazily; the file is closed when the epoch ends."""

    def __init__(self, path: str):
        self._path = path
        self._reader = None

    def __iter__(self) -> 'FileIterator':
        if self._reader is not None:
            self._reader.close()
        self._reader = open(self._path, 'r', encoding='utf-8')
        return self

    def __next__(self) -> str:
        assert self._reader is not None, 'Invalid Program State!'
        line = self._reader.readline()
        if line:
            return line
        self._reader.close()
        self._reader = None
        raise StopIteration('Finished.')
