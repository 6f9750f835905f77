"""TEST INFRASTRUCTURE ONLY: CPU restatements of the reference hot path (see the module headers).
Nothing under deepwalk-and-node2vec_b200/ may import this package."""
