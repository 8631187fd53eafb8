(* hnsw_b200_graph.ml — graph exchange between the reference's OCaml graphs and the GPU index, Hgraph.Stats,
   and the one-process multi-GPU index.

   * Of_ohnsw      walks an Ohnsw.Hgraph.t (lib/ohnsw.ml:306-351) into per-layer CSR and loads it on the GPU:
                   "a graph built by the reference, exported to the GPU layout" (the parity vehicle).
   * Export        the same for any Hnsw_algo.KNN_HGRAPH whose nodes are ints (lib/hnsw_algo.ml:944-958) — path A,
                   e.g. Hnsw.Ba's graph with ~id_base:1 (lib/hnsw.ml:313-325).
   * Downloaded    a GPU-built graph brought back as CSR arrays; it satisfies KNN_HGRAPH, so the reference's own
                   Knn functor (lib/hnsw_algo.ml:960-1011) runs over it, and it offers layer / adjacent /
                   Neighbours.iter with the shapes test/test.ml:51-66 uses on Ohnsw.Hgraph.
   * Stats         Hgraph.Stats (lib/hnsw.ml:353-375) computed on the device.
   * Multi         hnswb200_sharded_*: one call builds / queries all GPUs, as lib/ohnsw.ml:840-841,877 do one.

   NOT compiled in the build container (no OCaml toolchain there); the C side of every `external` below is
   type-checked against include/hnsw_b200.h by tests/test_ocaml_stubs_typecheck.py. *)

open Hnsw_b200

type i64vec = (int64, Bigarray.int64_elt, Bigarray.c_layout) Bigarray.Array1.t

external import_graph_ : index -> Lacaml.S.mat -> int -> int -> i64vec array -> i32vec array -> unit
  = "hb_import_graph_byte" "hb_import_graph"
external export_layer_nnz : index -> int -> int = "hb_export_layer_nnz"
external export_layer_ : index -> int -> int -> i64vec -> i32vec -> unit = "hb_export_layer"
external export_levels_ : index -> i32vec -> unit = "hb_export_levels"
external stats_ : index -> (int * int * int * float * int) array * int * int * float * float = "hb_stats"

(* one layer as CSR: row i = the list of node (i + id_base), head first (lib/ohnsw.ml:116-124) *)
type csr = { offsets : i64vec; nbrs : i32vec }

let csr_of_rows (rows : int list array) : csr =
  let n = Array.length rows in
  let offsets = Bigarray.Array1.create Bigarray.int64 Bigarray.c_layout (n + 1) in
  offsets.{0} <- 0L;
  Array.iteri (fun i r -> offsets.{i + 1} <- Int64.add offsets.{i} (Int64.of_int (List.length r))) rows;
  let nbrs = Bigarray.Array1.create Bigarray.int32 Bigarray.c_layout (max 1 (Int64.to_int offsets.{n})) in
  Array.iteri (fun i r -> List.iteri (fun j e -> nbrs.{Int64.to_int offsets.{i} + j} <- Int32.of_int e) r) rows;
  { offsets; nbrs }

let import (idx : index) (data : Lacaml.S.mat) ~id_base ~entry (layers : csr array) =
  import_graph_ idx data id_base entry (Array.map (fun c -> c.offsets) layers) (Array.map (fun c -> c.nbrs) layers)

(* ---- reference graph -> GPU ------------------------------------------------------------------------- *)

module Of_ohnsw = struct
  module O = Hnsw.Ohnsw
  let layer (h : _ O.Hgraph.t) l : csr =
    let g = O.Hgraph.layer h l in
    csr_of_rows (Array.init (O.Graph.num_nodes g) (fun i ->
        let row = ref [] in
        O.Neighbours.iter (O.Graph.adjacent g i) ~f:(fun e -> row := e :: !row);
        List.rev !row))
  (* load the reference-built graph `h` (over the vectors `data`) into a fresh GPU index *)
  let to_gpu ?(device = 0) (h : _ O.Hgraph.t) (data : Lacaml.S.mat) ~num_connections ~num_nodes_search_construction : index =
    let entry = match O.Hgraph.entry_point h with Some e -> e | None -> invalid_arg "knn: empty hgraph" in
    let idx = create (Lacaml.S.Mat.dim1 data) 0 num_connections num_nodes_search_construction 0 device in
    import idx data ~id_base:0 ~entry (Array.init (O.Hgraph.max_layer h + 1) (layer h));
    idx
end

module Export (H : Hnsw_algo.KNN_HGRAPH with type node = int) = struct
  let layer (h : H.t) l ~num_nodes ~id_base : csr =
    let g = H.layer h l in
    csr_of_rows (Array.init num_nodes (fun i ->
        List.rev (H.LayerGraph.Neighbours.fold (H.LayerGraph.adjacent g (i + id_base)) ~init:[] ~f:(fun acc e -> e :: acc))))
  let to_gpu ?(device = 0) ?(flavour = 1) (h : H.t) (data : Lacaml.S.mat) ~id_base ~num_neighbours ~num_neighbours_build : index =
    let idx = create (Lacaml.S.Mat.dim1 data) 0 num_neighbours num_neighbours_build 0 device in
    set_flavour idx flavour;
    import idx data ~id_base ~entry:(H.entry_point h)
      (Array.init (H.max_layer h + 1) (fun l -> layer h l ~num_nodes:(Lacaml.S.Mat.dim2 data) ~id_base));
    idx
end

(* ---- GPU graph -> OCaml ----------------------------------------------------------------------------- *)

module Downloaded = struct
  type node = int
  type value = Lacaml.S.vec
  let sexp_of_node = Base.Int.sexp_of_t
  let node_of_sexp = Base.Int.t_of_sexp
  let sexp_of_value (_ : value) = Base.Sexp.Atom "<vec>"
  let value_of_sexp _ = failwith "Downloaded.value_of_sexp"
  type t = { layers : csr array; entry : int; id_base : int; levels : i32vec; data : Lacaml.S.mat }

  module LayerGraph = struct
    type nonrec node = node
    let sexp_of_node = sexp_of_node
    let node_of_sexp = node_of_sexp
    type t = { csr : csr; base : int }
    let sexp_of_t (_ : t) = Base.Sexp.Atom "<layer>"
    let t_of_sexp _ = failwith "Downloaded.LayerGraph.t_of_sexp"
    let num_nodes g = Bigarray.Array1.dim g.csr.offsets - 1
    module Neighbours = struct
      type t = { nbrs : i32vec; first : int; last : int }
      let length n = n.last - n.first
      let fold n ~init ~f =
        let r = ref init in
        for i = n.first to n.last - 1 do r := f !r (Int32.to_int n.nbrs.{i}) done;
        !r
      let iter n ~f = for i = n.first to n.last - 1 do f (Int32.to_int n.nbrs.{i}) done
      let for_all n ~f = fold n ~init:true ~f:(fun acc e -> acc && f e)
      let is_empty n = n.last = n.first
    end
    let adjacent g node =
      let i = node - g.base in
      { Neighbours.nbrs = g.csr.nbrs; first = Int64.to_int g.csr.offsets.{i}; last = Int64.to_int g.csr.offsets.{i + 1} }
    (* the reference keeps a fresh N-sized array per query on this path (lib/hnsw_algo.ml:992) *)
    module Visited = struct
      type t_graph = t
      type t = { seen : Bytes.t; base : int; mutable count : int }
      let sexp_of_t (_ : t) = Base.Sexp.Atom "<visited>"
      let t_of_sexp _ = failwith "Downloaded.Visited.t_of_sexp"
      let create (g : t_graph) = { seen = Bytes.make (num_nodes g) '\000'; base = g.base; count = 0 }
      let mem v node = Bytes.get v.seen (node - v.base) <> '\000'
      let add v node = if not (mem v node) then (Bytes.set v.seen (node - v.base) '\001'; v.count <- v.count + 1); v
      let length v = v.count
      let clear v = Bytes.fill v.seen 0 (Bytes.length v.seen) '\000'; v.count <- 0; v
    end
  end

  let layer h l = { LayerGraph.csr = h.layers.(l); base = h.id_base }
  let max_layer h = Array.length h.layers - 1
  let entry_point h = h.entry
  let num_nodes h = Lacaml.S.Mat.dim2 h.data
  let value h node = Lacaml.S.Mat.col h.data (node - h.id_base + 1)
  let level h node = Int32.to_int h.levels.{node - h.id_base}

  (* bring the graph of `idx` (built over `data`) back to the host; ~id_base:1 numbers nodes as Hnsw.Ba does *)
  let download ?(id_base = 0) (idx : index) (data : Lacaml.S.mat) : t =
    let (n, max_layer, entry) = info idx in
    let layers = Array.init (max_layer + 1) (fun l ->
        let offsets = Bigarray.Array1.create Bigarray.int64 Bigarray.c_layout (n + 1) in
        let nbrs = Bigarray.Array1.create Bigarray.int32 Bigarray.c_layout (max 1 (export_layer_nnz idx l)) in
        export_layer_ idx l id_base offsets nbrs;
        { offsets; nbrs }) in
    let levels = Bigarray.Array1.create Bigarray.int32 Bigarray.c_layout n in
    export_levels_ idx levels;
    { layers; entry = entry + id_base; id_base; levels; data }
end

(* Hnsw_algo.Knn over a GPU-built graph: module K = Hnsw_algo.Knn (Downloaded) (VisitMe) (Nearest) (Distance) *)
module _ : Hnsw_algo.KNN_HGRAPH with type node = int = Downloaded

(* ---- Hgraph.Stats (lib/hnsw.ml:353-375) ------------------------------------------------------------- *)

module Stats = struct
  type mima = { min : int; max : int; mean : float; isolated : int }   (* the reference lists the isolated nodes; here their number *)
  type t = { num_nodes : int; layer_sizes : int array; layer_connectivity : mima array;
             search_distance_computations : int; build_distance_computations : int }   (* lib/hnsw.ml:732-751 *)
  let of_tuple num_nodes (layers, sd, bd, _, _) =
    { num_nodes;
      layer_sizes = Array.map (fun (nodes, _, _, _, _) -> nodes) layers;
      layer_connectivity = Array.map (fun (_, min, max, mean, isolated) -> { min; max; mean; isolated }) layers;
      search_distance_computations = sd; build_distance_computations = bd }
  let compute (idx : index) = let (n, _, _) = info idx in of_tuple n (stats_ idx)
end

(* ---- one process, every GPU ------------------------------------------------------------------------- *)

module Multi = struct
  type t
  external create_ : int -> int -> int -> int -> int -> int array -> t = "hb_sharded_create_byte" "hb_sharded_create"
  external close : t -> unit = "hb_sharded_close"
  external set_flavour : t -> int -> unit = "hb_sharded_set_flavour"
  external build_ : t -> Lacaml.S.mat -> i32vec -> unit = "hb_sharded_build"
  external search_ : t -> Lacaml.S.mat -> int -> int -> i32mat -> Lacaml.S.mat -> unit = "hb_sharded_search_byte" "hb_sharded_search"
  external stats_ : t -> (int * int * int * float * int) array * int * int * float * float = "hb_sharded_stats"
  external num_nodes : t -> int = "hb_sharded_num_nodes"

  (* Ohnsw.build_batch_bigarray (lib/ohnsw.ml:840-857): rows cut into one contiguous shard per device *)
  let build_batch_bigarray ?(seed = 0) ~devices (distance : distance) (batch : Lacaml.S.mat)
      ~num_connections ~num_nodes_search_construction : t =
    let h = create_ (Lacaml.S.Mat.dim1 batch) (distance_tag distance) num_connections num_nodes_search_construction seed devices in
    build_ h batch no_levels;
    h

  (* Ohnsw.knn_batch_bigarray (lib/ohnsw.ml:877-897): global ids, ascending by (distance, id) over all shards *)
  let knn_batch_bigarray ?ef (h : t) ~k (batch : Lacaml.S.mat) =
    let nq = Lacaml.S.Mat.dim2 batch in
    let distances = Lacaml.S.Mat.create k nq in
    let ids32 = Bigarray.Array2.create Bigarray.int32 Bigarray.c_layout nq k in
    search_ h batch k (match ef with Some e -> e | None -> k) ids32 distances;
    Array.init nq (fun j -> Array.init k (fun i -> Int32.to_int ids32.{j, i})), distances

  let stats h = Stats.of_tuple (num_nodes h) (stats_ h)
end
