"""`merge_window_matches_unique_ref` with the reference's signature (src/helpers.py:692-815).

Host-side, user-invoked post-processing on at most N rows (outside the GPU hot path):
concatenate per-window matches, de-duplicate identical (aligned, ref) pairs preferring non-violating rows then
smaller `window_id`, and keep a maximum-cardinality one-to-one matching (Hopcroft-Karp)."""
from __future__ import annotations

import pandas as pd


def merge_window_matches_unique_ref(matches_list, cell_id_col="Cell_Num_Old"):
    if not matches_list:
        return pd.DataFrame()
    import networkx as nx
    merged = pd.concat(matches_list, ignore_index=True)
    a_col, r_col = f"Aligned_{cell_id_col}", f"Ref_{cell_id_col}"
    missing = [c for c in ["window_id", a_col, r_col, "X", "Y", "filtered_violation"] if c not in merged.columns]
    if missing:
        raise ValueError(f"Missing required columns in matches: {missing}")
    merged["filtered_violation"] = merged["filtered_violation"].fillna(True).astype(bool)
    merged = merged.sort_values(by=["filtered_violation", "window_id"], ascending=[True, True], kind="mergesort")
    merged = merged.drop_duplicates(subset=[a_col, r_col], keep="first")
    a_vals, r_vals = merged[a_col].values, merged[r_col].values
    ua, ur = sorted(pd.unique(a_vals)), sorted(pd.unique(r_vals))
    edge_row = {(a, b): k for k, (a, b) in enumerate(zip(a_vals, r_vals))}
    g = nx.Graph()
    a_nodes = [f"align_{a}" for a in ua]
    g.add_nodes_from(a_nodes, bipartite=0)
    g.add_nodes_from([f"ref_{b}" for b in ur], bipartite=1)
    adj = {a: [] for a in ua}
    for a, b in zip(a_vals, r_vals):
        adj[a].append(b)
    for a in ua:                       # same edge insertion order as the reference (per aligned id, row order)
        for b in adj[a]:
            g.add_edge(f"align_{a}", f"ref_{b}")
    matching = nx.bipartite.hopcroft_karp_matching(g, top_nodes=set(a_nodes))
    a_of = {f"align_{a}": a for a in ua}
    r_of = {f"ref_{b}": b for b in ur}
    rows = []
    for lab in a_nodes:
        m = matching.get(lab)
        if m:
            k = edge_row.get((a_of[lab], r_of[m]))
            if k is not None:
                rows.append(k)
    return merged.iloc[rows].copy().reset_index(drop=True)
