"""Host-side logic of the product that needs no GPU: the SentencePiece model loader's table building
(csrc/sp_model.cc, through xllm_tokenizer_probe), error codes, and the workload generator."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(__file__)
MODEL_DIR = os.path.join(HERE, "golden", "sp_bpe_8k")


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build()


def test_tokenizer_probe_fixture():
    from xllm_service_b200 import _lib
    info = _lib.tokenizer_probe(MODEL_DIR)
    assert info["n_pieces"] == 8000 and info["byte_fallback"] == 1 and info["unk_id"] == 0
    assert info["split_mode"] == 1          # no piece holds U+2581 past its first char: words split exactly
    assert info["n_symbols"] >= info["n_pieces"]
    assert info["n_pairs"] > 8000 and info["n_pair_slots"] >= 4 * info["n_pairs"]   # load factor <= 0.25
    assert info["trie_units"] > 0 and 3 <= info["max_unit_out"] <= 64
    # the file form and the directory form resolve to the same model
    assert _lib.tokenizer_probe(os.path.join(MODEL_DIR, "tokenizer.model")) == info


def test_tokenizer_probe_errors(tmp_path):
    import xllm_service_b200 as x
    from xllm_service_b200 import _lib
    with pytest.raises(x.IngestError) as e:
        _lib.tokenizer_probe(str(tmp_path / "nope"))
    assert e.value.code == -3  # XLLM_ERR_IO
    (tmp_path / "tokenizer.model").write_bytes(b"not a protobuf \xff\xff\xff")
    with pytest.raises(x.IngestError) as e:
        _lib.tokenizer_probe(str(tmp_path))
    assert e.value.code in (-4, -5)


def test_unigram_model_is_rejected_not_mistokenised(tmp_path):
    spm = pytest.importorskip("sentencepiece")
    import io
    import xllm_service_b200 as x
    from xllm_service_b200 import _lib, workload
    model = io.BytesIO()
    spm.SentencePieceTrainer.train(sentence_iterator=iter(workload.sentences(2000, seed=1)), model_writer=model,
                                   model_type="unigram", vocab_size=400, minloglevel=2)
    (tmp_path / "tokenizer.model").write_bytes(model.getvalue())
    with pytest.raises(x.IngestError) as e:
        _lib.tokenizer_probe(str(tmp_path))
    assert e.value.code == -5  # XLLM_ERR_UNSUPPORTED


def test_workload_exact_tokens_and_shared_prefixes(oracle):
    from xllm_service_b200 import workload
    sp = oracle.SentencePieceOracle(MODEL_DIR)
    vocab = workload.make_vocabulary()
    assert len(vocab) == 20000 and len(set(vocab)) == 20000 and all(2 <= len(w) <= 9 for w in vocab)
    wb = workload.pack_prompts(vocab)
    _, wcnt = sp.encode_batch(wb.text, wb.offsets, 32, n_threads=4)
    T = 512
    b, meta = workload.make_prompts_exact_tokens(
        200, T, wcnt, seed=3, shared_prefix=dict(n_prefixes=6, frac=0.8, min_blocks=1, max_blocks=3, block_tokens=128))
    b2, _ = workload.make_prompts_exact_tokens(
        200, T, wcnt, seed=3, shared_prefix=dict(n_prefixes=6, frac=0.8, min_blocks=1, max_blocks=3, block_tokens=128))
    assert (b.text == b2.text).all() and (b.offsets == b2.offsets).all()   # deterministic
    ids, n = sp.encode_batch(b.text, b.offsets, T + 8, n_threads=4)
    assert (n == T).all()
    pid = meta["prefix_id"]
    assert 0.6 < (pid >= 0).mean() < 0.95
    for j in np.unique(pid[pid >= 0]):
        rows = np.nonzero(pid == j)[0]
        L = int(meta["prefix_blocks"][rows[0]]) * 128
        assert (ids[rows, :L] == ids[rows[0], :L]).all()


def test_factory_selects_tiktoken_by_tokenizer_class(tmp_path):
    """TokenizerFactory::create_tokenizer (tokenizer_factory.cpp:9-32): tokenizer_class == TikTokenTokenizer
    -> tiktoken tables (bytes are symbols, no pre-split, no normaliser); otherwise SentencePiece."""
    import shutil
    from xllm_service_b200 import _lib
    info = _lib.tokenizer_probe(os.path.join(HERE, "golden", "tiktoken_1k"))
    assert info["split_mode"] == 0 and info["trie_units"] == 0 and info["byte_fallback"] == 0
    assert info["n_pieces"] == 1453            # 256 bytes + 1200 merges - 3 bytes dropped on purpose
    assert info["n_symbols"] == 1456           # the 3 rank-less bytes are still symbols (they emit nothing)
    assert info["n_pairs"] >= 1200
    # the same vocabulary file without the tokenizer_class hint is NOT silently read as tiktoken
    d = tmp_path / "plain"
    d.mkdir()
    shutil.copy(os.path.join(HERE, "golden", "tiktoken_1k", "tokenizer.model"), d / "tokenizer.model")
    import xllm_service_b200 as x
    with pytest.raises(x.IngestError):
        _lib.tokenizer_probe(str(d))
