"""Model-level chunk loop of the LTM path.

In the reference the loop over the chunks of a video lives in the eval scripts: they call the video Q-former once per
chunk with `new_video=(chunk == 0)` (infty-Video-LLaMA/InfVideoLLaMA/models/infinityqa.py:280-344 `encode_video`,
infty-VideoChat2/models/videochat2_it_mistral.py:181-253 `encode_img`) and keep a running mean of the chunk embeddings
(eval_code/eval/run_inference_inf_video_llama_nextqa.py:194).  Inside the Q-former every LTM layer (2 in
Video-LLaMA, 6 in VideoChat2) is handed the SAME `encoder_hidden_states` and pools its frames again
(long_term_attention_gibbs.py:304).

`consolidate_video` is that loop for a batch of independent videos and all LTM layers at once:
  * the frames of a chunk are pooled ONCE (25 MB per video at the NExT-QA shape) and shared by every layer;
  * the pooling of chunk c+1 runs on a side stream while the layers consolidate chunk c;
  * chunks of a video stay strictly sequential, videos and layers are independent.
"""
import torch

from .batched import BatchedRectLTM


@torch.no_grad()
def consolidate_video(engines, chunks, queries, uniforms=None, new_doc=True, reduce=None):
    """engines:  one `BatchedRectLTM` per LTM layer (same device, same chunk geometry).
    chunks:   C tensors [Bv, L*T, e] -- the `encoder_hidden_states` of each chunk, in order.
    queries:  `queries[layer][c]` tensors [Bv, Q, D], or a callable `queries(layer, c, prev)` returning that tensor
              (`prev` = context of the previous layer for this chunk, None for the first layer): in the Q-former the
              queries of a layer depend on the output of the layer before it.
    uniforms: `uniforms[layer][c]` float64 [Bv, 512] for c >= 1 (sticky engines); None entries allowed at c == 0.
    new_doc:  whether chunk 0 starts a new document for every video.
    reduce:   None -> `ctx[layer][c]`;  "mean" -> running mean over the chunks, `ctx[layer]` [Bv, Q, D].
    """
    if not engines or any(not isinstance(e, BatchedRectLTM) for e in engines):
        raise ValueError("consolidate_video takes a list of BatchedRectLTM engines (one per LTM layer)")
    dev = engines[0].device
    if any(e.device != dev or e.T != engines[0].T or e.e != engines[0].e for e in engines):
        raise ValueError("all layers must live on one device and share the chunk geometry")
    if reduce not in (None, "mean"):
        raise ValueError("reduce must be None or 'mean'")
    C = len(chunks)
    pool_eng = engines[0]
    with torch.cuda.device(dev):
        main = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(main)                                    # the chunks were produced on the caller's stream
        pooled, ready = [None] * C, [None] * C

        def pool_ahead(c):
            with torch.cuda.stream(side):
                pooled[c] = pool_eng.pool(chunks[c])
                ready[c] = torch.cuda.Event()
                ready[c].record(side)

        pool_ahead(0)
        out = [[] for _ in engines]
        acc = [None] * len(engines)
        for c in range(C):
            main.wait_event(ready[c])
            pooled[c].record_stream(main)
            if c + 1 < C:
                pool_ahead(c + 1)
            prev = None
            for li, eng in enumerate(engines):
                q = queries(li, c, prev) if callable(queries) else queries[li][c]
                u = None if uniforms is None else uniforms[li][c]
                ctx = eng.step(chunks[c], q, u if c > 0 or not new_doc else None, new_doc=(new_doc and c == 0),
                               pooled=pooled[c], beside_pooling=(c + 1 < C))
                prev = ctx
                if reduce == "mean":
                    acc[li] = ctx.clone() if acc[li] is None else acc[li].add_(ctx)
                else:
                    out[li].append(ctx)
            pooled[c] = None
        if reduce == "mean":
            return [a / C for a in acc]
        return out
