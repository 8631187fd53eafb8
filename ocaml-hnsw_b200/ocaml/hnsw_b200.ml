(* hnsw_b200.ml — the reference's build / k-NN entry points backed by libhnsw_b200.so.

   Same labelled signatures as lib/ohnsw.ml (build_batch_bigarray :840, insert :766, knn :859,
   knn_batch_bigarray :877) and as Hnsw.Ba (lib/hnsw.ml:753-777), so benchmark/benchmark.ml runs
   unchanged after `module Ohnsw = Hnsw_b200.Ohnsw`.  The one narrowing: `'a distance` is a tag,
   not a closure.

   NOT compiled in the build container (no OCaml toolchain there). *)

type index
type distance = L2 | Angular | Ip
let distance_tag = function L2 -> 0 | Angular -> 1 | Ip -> 2

type i32mat = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array2.t
type i32vec = (int32, Bigarray.int32_elt, Bigarray.c_layout) Bigarray.Array1.t

external create : int -> int -> int -> int -> int -> int -> index = "hb_create_byte" "hb_create"
external close : index -> unit = "hb_close"
external set_flavour : index -> int -> unit = "hb_set_flavour"
external build_ : index -> Lacaml.S.mat -> i32vec -> unit = "hb_build"
external insert_ : index -> Lacaml.S.mat -> i32vec -> unit = "hb_insert"
external search_ : index -> Lacaml.S.mat -> int -> int -> i32mat -> Lacaml.S.mat -> unit = "hb_search_byte" "hb_search"
external info : index -> int * int * int = "hb_info"
external params : index -> int * int = "hb_params"            (* num_connections, num_nodes_search_construction *)
external bruteforce_ : Lacaml.S.mat -> Lacaml.S.mat -> int -> Lacaml.S.mat -> unit = "hb_bruteforce"
(* hnswb200_host_register on a Bigarray payload (off-heap, never moved by the GC).  knn_batch_bigarray on pinned
   query / result Bigarrays copies nothing: the search kernel reads each query from the Bigarray and stores each result
   row into it over PCIe while it runs (include/hnsw_b200.h). *)
external pin : ('a, 'b, 'c) Bigarray.Genarray.t -> unit = "hb_pin"
external unpin : ('a, 'b, 'c) Bigarray.Genarray.t -> unit = "hb_unpin"

let no_levels : i32vec = Bigarray.Array1.create Bigarray.int32 Bigarray.c_layout 0

module Ohnsw = struct
  module Hgraph = struct
    type _ t = index
    let num_nodes h = let (n, _, _) = info h in n                          (* lib/ohnsw.ml:335 *)
    let max_layer h = let (_, l, _) = info h in l                          (* :346 *)
    let entry_point h = let (_, _, e) = info h in if e < 0 then None else Some e   (* :340 *)
  end
  module Visited = struct
    type t = unit                      (* visited sets live in GPU shared memory, per query *)
    let create (_ : int) = ()
  end
  (* lib/ohnsw.ml:6-12 *)
  module HeapElt = struct
    type t = { node : int; distance : float }
  end
  (* The result queue of Ohnsw.knn (lib/ohnsw.ml:404-416): abstract, popped in ascending distance.  Only the
     operations a caller of knn uses on its result are offered (pop_min, iter, copy). *)
  module MinQueue : sig
    type 'a t
    type element = HeapElt.t
    val of_sorted : element list -> 'a t
    val pop_min : 'a t -> element option
    val iter : 'a t -> f:(element -> unit) -> unit
    val copy : 'a t -> 'a t
  end = struct
    type 'a t = { mutable rest : HeapElt.t list }
    type element = HeapElt.t
    let of_sorted rest = { rest }
    let pop_min q = match q.rest with [] -> None | e :: tl -> q.rest <- tl; Some e
    let iter q ~f = List.iter f q.rest
    let copy q = { rest = q.rest }
  end

  let distance_l2 = L2                                                     (* lib/ohnsw.ml:899 *)

  (* lib/ohnsw.ml:840-857 *)
  let build_batch_bigarray ?(seed = 0) ?(device = 0) (distance : distance) (batch : Lacaml.S.mat)
      ~num_connections ~num_nodes_search_construction : _ Hgraph.t =
    let h = create (Lacaml.S.Mat.dim1 batch) (distance_tag distance) num_connections
        num_nodes_search_construction seed device in
    build_ h batch no_levels;
    h

  (* lib/ohnsw.ml:766-837; M, efC and level_mult (= 1 / ln M, :844) were fixed at creation: the reference passes
     the same values on every call, other values are rejected *)
  let insert (h : _ Hgraph.t) (target : Lacaml.S.vec) ~num_connections ~num_nodes_search_construction
      (level_mult : float) (_ : Visited.t) =
    let (m_, efc_) = params h in
    if num_connections <> m_ then invalid_arg "insert: num_connections differs from the index's";
    if num_nodes_search_construction <> efc_ then invalid_arg "insert: num_nodes_search_construction differs from the index's";
    if Float.abs (level_mult -. 1. /. log (float_of_int m_)) > 1e-9 *. level_mult then
      invalid_arg "insert: level_mult differs from 1 / ln num_connections";
    let m = Lacaml.S.Mat.of_col_vecs [| target |] in
    insert_ h m no_levels

  (* lib/ohnsw.ml:877-897: ids as int array array (nq x k, -1 padded), distances k x nq (NaN padded) *)
  let knn_batch_bigarray ?ef (h : _ Hgraph.t) ~k (batch : Lacaml.S.mat) =
    let nq = Lacaml.S.Mat.dim2 batch in
    let distances = Lacaml.S.Mat.create k nq in
    let ids32 = Bigarray.Array2.create Bigarray.int32 Bigarray.c_layout nq k in
    search_ h batch k (match ef with Some e -> e | None -> k) ids32 distances;
    let ids = Array.init nq (fun j -> Array.init k (fun i -> Int32.to_int ids32.{j, i})) in
    ids, distances

  (* lib/ohnsw.ml:859-875: the result as a MinQueue.t, popped nearest first *)
  let knn (h : _ Hgraph.t) (_ : Visited.t) ~k (target : Lacaml.S.vec) : _ MinQueue.t =
    let ids, d = knn_batch_bigarray h ~k (Lacaml.S.Mat.of_col_vecs [| target |]) in
    MinQueue.of_sorted
      (List.filter (fun (e : HeapElt.t) -> e.node >= 0)
         (List.init k (fun i -> { HeapElt.node = ids.(0).(i); distance = d.{i + 1, 1} })))
end

(* Hnsw.Ba = MakeBatch(EuclideanBa), lib/hnsw.ml:729-778, 817-819 *)
module Ba = struct
  type t = index
  type value = Lacaml.S.vec
  let build ?(seed = 0) ?(device = 0) ~num_neighbours ~num_neighbours_build (data : Lacaml.S.mat) : t =
    let h = create (Lacaml.S.Mat.dim1 data) 0 num_neighbours num_neighbours_build seed device in
    set_flavour h 1;
    build_ h data no_levels;
    h
  let search (h : t) (batch : Lacaml.S.mat) ~num_neighbours_search ~num_neighbours =
    let nq = Lacaml.S.Mat.dim2 batch in
    let distances = Lacaml.S.Mat.create num_neighbours nq in
    let ids32 = Bigarray.Array2.create Bigarray.int32 Bigarray.c_layout nq num_neighbours in
    search_ h batch num_neighbours (max num_neighbours_search num_neighbours) ids32 distances;
    ids32, distances
  (* lib/hnsw.ml:769-777: distances only, +inf padded (:770) *)
  let knn_batch (h : t) (batch : Lacaml.S.mat) ~num_neighbours_search ~num_neighbours : Lacaml.S.mat =
    snd (search h batch ~num_neighbours_search ~num_neighbours)
  (* lib/hnsw.ml:763-767: (node, distance) pairs, nearest first; Hnsw.Ba numbers nodes from 1 (:313-325).
     A maintainer maps the pair onto Hnsw_algo.value_distance (lib/hnsw_algo.ml:85). *)
  let knn (h : t) (point : value) ~num_neighbours_search ~num_neighbours =
    let ids, d = search h (Lacaml.S.Mat.of_col_vecs [| point |]) ~num_neighbours_search ~num_neighbours in
    List.filter (fun (i, _) -> i >= 1)
      (List.init num_neighbours (fun i -> Int32.to_int ids.{0, i} + 1, d.{i + 1, 1}))
end

(* benchmark/dataset.ml:15-30 *)
let brute_force_knn_l2 (train : Lacaml.S.mat) (test : Lacaml.S.mat) k : Lacaml.S.mat =
  let d = Lacaml.S.Mat.create k (Lacaml.S.Mat.dim2 test) in
  bruteforce_ train test k d;
  d
