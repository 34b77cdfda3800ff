;;; bridgemc-cffi.lisp -- CFFI shim that puts libbridgemc.so behind bridge-mc's Lisp entry points
;;; (see INTEGRATION.md).  Load AFTER bridge.cl / bidding.cl / compscore.cl:
;;;
;;;     (load "bridge.cl") (load "bidding.cl") (load "compscore.cl") (load "bridgemc-cffi.lisp")
;;;
;;; It re-points, under their own names and signatures,
;;;     build  sim  assess-hand  test-bid  choose-bid  contract-score  base-score  match-points
;;; at the library and adds helpers that feed device tallies into hist-base / histogram.  rest.cl and
;;; the frontend are untouched.  NOT EXECUTED in the build image (no Lisp implementation there): every
;;; foreign call made here is mirrored call for call by bridge_mc_b200/ (ctypes) and exercised by
;;; tests/test_gpu_shim_mirror.py with the same argument conventions.

(ql:quickload :cffi)

(cffi:define-foreign-library libbridgemc (t (:default "libbridgemc")))
(cffi:use-foreign-library libbridgemc)

;;; ---- foreign types (include/bridgemc.h) ------------------------------------------------------
(cffi:defcstruct bmc-record (lo :uint64) (hi :uint64))
(cffi:defcstruct bmc-seat-spec
  (hcp :int8 :count 2) (len :int8 :count 8) (suit-hcp :int8 :count 8) (losers :int8 :count 2)
  (fixed :uint8) (reserved :uint8 :count 3) (fixed-mask :uint64))
(cffi:defcstruct bmc-contract (level :int8) (suit :int8) (declarer :int8) (dbl :int8))
(cffi:defcstruct bmc-stats
  (raw :uint64) (voided :uint64) (accepted :uint64) (stored :uint64)
  (waves :uint32) (launches :uint32) (kernel-ms :float) (total-ms :float))
(cffi:defcstruct bmc-seat-features
  (len :uint8 :count 4) (suit-hcp :uint8 :count 4) (hcp :uint8) (losers :uint8) (balanced :uint8) (bid :int8))
(cffi:defcstruct bmc-measure (kind :uint8) (seat :uint8) (suit :uint8) (reserved :uint8))
;; bmc_tallies as plain 64-bit words: 4 counters | hcp[4][40] | len[4][4][14] | shape[4][40] |
;; tricks[8][14] | imps[8][49] | imp_dropped[8]
(defconstant +tally-words+ (+ 4 160 224 160 112 392 8))
(defconstant +tally-hcp+ 4)
(defconstant +tally-len+ (+ 4 160))
(defconstant +tally-shape+ (+ 4 160 224))
(defconstant +tally-tricks+ (+ 4 160 224 160))
(defconstant +tally-imps+ (+ 4 160 224 160 112))

(cffi:defcfun "bmc_init" :int (device :int) (out :pointer))
(cffi:defcfun "bmc_shutdown" :void (ctx :pointer))
(cffi:defcfun "bmc_last_error" :string)
(cffi:defcfun "bmc_constraints_create" :int (specs :pointer) (rules :pointer) (out :pointer))
(cffi:defcfun "bmc_constraints_destroy" :void (c :pointer))
(cffi:defcfun "bmc_constraints_set_reference_order" :int (c :pointer) (n-defs :int))
(cffi:defcfun "bmc_constraints_set_mode" :int (c :pointer) (mode :int))
(cffi:defcfun "bmc_constraints_set_jit" :int (c :pointer) (on :int))
(cffi:defcfun "bmc_constraints_set_scoring" :int
  (c :pointer) (contracts :pointer) (vulnerable :pointer) (n-contracts :uint32) (pairs :pointer) (n-pairs :uint32))
;; what determines the deal behind a Philox index besides the seed (log them with a reproducible job)
(cffi:defcfun "bmc_constraints_mode" :int (c :pointer))
(cffi:defcfun "bmc_constraints_primary" :int (c :pointer))
(cffi:defcfun "bmc_constraints_primary_hcp" :int (c :pointer) (lo :pointer) (hi :pointer))
(cffi:defcfun "bmc_scheme_create" :int (table :string) (out :pointer))
(cffi:defcfun "bmc_scheme_destroy" :void (s :pointer))
(cffi:defcfun "bmc_sample" :int
  (ctx :pointer) (cons :pointer) (seed :uint64) (first-index :uint64) (n-raw :uint64)
  (n-accept :uint64) (wave :uint64) (tally-mask :uint32)
  (records :pointer) (cap :uint64) (tallies :pointer) (stats :pointer))
(cffi:defcfun "bmc_sample_multi" :int
  (devices :pointer) (n-devices :int) (cons :pointer) (seed :uint64) (first-index :uint64) (n-raw :uint64)
  (n-accept :uint64) (wave :uint64) (tally-mask :uint32)
  (records :pointer) (cap :uint64) (tallies :pointer) (stats :pointer))
(cffi:defcfun "bmc_filter" :int
  (ctx :pointer) (cons :pointer) (scheme :pointer) (records :pointer) (n :uint64)
  (accept :pointer) (feats :pointer))
(cffi:defcfun "bmc_score" :int
  (ctx :pointer) (contracts :pointer) (vulnerable :pointer) (n-contracts :uint32)
  (tricks :pointer) (n :uint64) (scores :pointer))
(cffi:defcfun "bmc_match_points" :int
  (ctx :pointer) (a :pointer) (b :pointer) (n :uint64) (imps :pointer) (status :pointer))
(cffi:defcfun "bmc_hist_base" :int
  (ctx :pointer) (records :pointer) (n :uint64) (measures :pointer) (k :uint32)
  (tuples :pointer) (counts :pointer) (cap :uint64) (n-unique :pointer))

;;; ---- errors, context, devices ------------------------------------------------------------------
(defun bmc-check (rc)
  (cond ((= rc 0) t)
        ((= rc -3) (error 'bad-range :msg (bmc-last-error) :down nil :up nil))   ; bridge.cl:1109-1123
        (t (error "libbridgemc: ~A (~A)" (bmc-last-error) rc))))

(defmacro with-gpu-call (&body body)
  "SBCL enables floating-point traps that the CUDA runtime trips: mask them around every foreign call."
  `(sb-int:with-float-traps-masked (:invalid :divide-by-zero :overflow) ,@body))

(defvar *bmc-devices* '(0)
  "CUDA devices sampling jobs run on.  More than one: build / sim go through bmc_sample_multi (one
host thread per device inside the library, tallies summed there); scoring and filtering use the first.")
(defvar *bmc-ctx* nil)
(defvar *bmc-lock* (bt:make-lock "bridgemc"))

(defun bmc-ctx ()
  "One context shared by all threads (hunchentoot workers, dealgen<..> / sim{..} threads): the library
serialises calls on a context; create more with bmc-init for concurrency."
  (or *bmc-ctx*
      (bt:with-lock-held (*bmc-lock*)
        (or *bmc-ctx*
            (cffi:with-foreign-object (out :pointer)
              (with-gpu-call (bmc-check (bmc-init (first *bmc-devices*) out)))
              (setf *bmc-ctx* (cffi:mem-ref out :pointer)))))))

(defun bmc-close ()
  (bt:with-lock-held (*bmc-lock*)
    (when *bmc-ctx* (bmc-shutdown *bmc-ctx*) (setf *bmc-ctx* nil))))

;;; ---- records <-> hand objects (bridge.cl:1622-1633) ------------------------------------------------
(defun record->hands (lo hi)
  (loop for seat from 0 to 3
        collect (mkhand
                 (loop for suit from 0 to 3
                       collect (loop for rank from 0 to 12
                                     for bit = (+ (* 16 suit) rank)
                                     when (and (eq (logbitp bit lo) (logbitp 0 seat))
                                               (eq (logbitp bit hi) (logbitp 1 seat)))
                                       collect rank)))))

(defun hand->lane-mask (hand)
  (loop for suit from 0 to 3
        sum (loop for rank in (nth suit (suits hand)) sum (ash 1 (+ (* 16 suit) rank)))))

(defun hand->record (hand)
  "A deal whose seat 0 is HAND (the other 39 cards go to seats 1..3 in card order): what the per-hand
entry points (assess-hand, test-bid, choose-bid) hand to bmc_filter.  Returns (values lo hi)."
  (let ((mine (hand->lane-mask hand)) (lo 0) (hi 0) (n 0))
    (loop for suit from 0 to 3
          do (loop for rank from 0 to 12
                   for bit = (+ (* 16 suit) rank)
                   unless (logbitp bit mine)
                     do (let ((seat (+ 1 (floor n 13))))
                          (when (logbitp 0 seat) (setf lo (logior lo (ash 1 bit))))
                          (when (logbitp 1 seat) (setf hi (logior hi (ash 1 bit))))
                          (incf n))))
    (values lo hi)))

;;; ---- range defs -> bmc_seat_spec ---------------------------------------------------------------------
(defun fill-range (ptr slot index spec)
  "spec: number | (lo hi ...) | nil ; -1 = nil"
  (let ((lo (cond ((numberp spec) spec) ((listp spec) (first spec))))
        (hi (cond ((numberp spec) spec) ((listp spec) (second spec)))))
    (setf (cffi:mem-aref (cffi:foreign-slot-pointer ptr '(:struct bmc-seat-spec) slot) :int8 (* 2 index)) (or lo -1)
          (cffi:mem-aref (cffi:foreign-slot-pointer ptr '(:struct bmc-seat-spec) slot) :int8 (+ 1 (* 2 index))) (or hi -1))))

(defun clear-spec (p)
  (loop for i from 0 below (cffi:foreign-type-size '(:struct bmc-seat-spec))
        do (setf (cffi:mem-aref p :uint8 i) 0))
  (fill-range p 'hcp 0 nil) (fill-range p 'losers 0 nil)
  (dotimes (s 4) (fill-range p 'len s nil) (fill-range p 'suit-hcp s nil)))

;;; ---- class deal: `build` keeps its signature (bridge.cl:1292) ----------------------------------------
;;; rest.cl makes plain `deal` objects, so the state of the GPU generator hangs off the object in a weak
;;; table instead of new slots, and `build` is redefined for class `deal` itself.
(defstruct gpu-gen handle (buffer nil) (next-index 0) (seed (random (ash 1 62))))
(defvar *gpu-gens* (make-hash-table :test 'eq :weakness :key :synchronized t))

(defun make-deal-handle (const-hands defs)
  (cffi:with-foreign-objects ((specs '(:struct bmc-seat-spec) 4) (out :pointer))
    (dotimes (k 4) (clear-spec (cffi:mem-aptr specs '(:struct bmc-seat-spec) k)))
    (loop for hand in const-hands
          for k from 0
          for p = (cffi:mem-aptr specs '(:struct bmc-seat-spec) k)
          do (setf (cffi:foreign-slot-value p '(:struct bmc-seat-spec) 'fixed) 1
                   (cffi:foreign-slot-value p '(:struct bmc-seat-spec) 'fixed-mask)
                   ;; const hands are "pretty" suit lists of rank symbols: the reference's own conversion chain
                   ;; (bridge.cl:1316: unhand -> hand -> mkhand) gives the hand object
                   (hand->lane-mask (mkhand (hand (unhand hand))))))
    (loop for def in defs
          for k from (length const-hands)
          for p = (cffi:mem-aptr specs '(:struct bmc-seat-spec) k)
          do (destructuring-bind (&key hcp c d h s &allow-other-keys) def
               (fill-range p 'hcp 0 hcp)
               (loop for spec in (list c d h s) for i from 0
                     do (fill-range p 'len i spec)
                        (fill-range p 'suit-hcp i (and (listp spec) (third spec))))))
    (with-gpu-call (bmc-check (bmc-constraints-create specs (cffi:null-pointer) out)))
    (let ((handle (cffi:mem-ref out :pointer)))
      ;; keep `build`'s own seat-by-seat distribution (and quirks); no def still samples one seat
      (with-gpu-call (bmc-check (bmc-constraints-set-reference-order handle (max 1 (length defs)))))
      handle)))

(defmethod gpu-gen ((this deal))
  (or (gethash this *gpu-gens*)
      (with-slots (defs const-hands) this
        (let* ((handle (make-deal-handle const-hands defs))
               (gen (make-gpu-gen :handle handle)))
          ;; the library object lives as long as the deal does (the closure must not capture `this`)
          (sb-ext:finalize this (lambda () (bmc-constraints-destroy handle)) :dont-save t)
          (setf (gethash this *gpu-gens*) gen)))))

(defun sample-records (handle seed first-index n &key (tally-mask 0) tallies)
  "N accepted deals from one library call.  Returns (values list-of-(lo hi) indices-used)."
  (let ((cap (+ n 16384)))
    (cffi:with-foreign-objects ((recs '(:struct bmc-record) cap) (stats '(:struct bmc-stats)))
      (with-gpu-call
        (bmc-check
         (if (rest *bmc-devices*)
             (cffi:with-foreign-object (devs :int (length *bmc-devices*))
               (loop for d in *bmc-devices* for i from 0 do (setf (cffi:mem-aref devs :int i) d))
               (bmc-sample-multi devs (length *bmc-devices*) handle seed first-index 0 n 16384 tally-mask
                                 recs cap (or tallies (cffi:null-pointer)) stats))
             (bmc-sample (bmc-ctx) handle seed first-index 0 n 16384 tally-mask
                         recs cap (or tallies (cffi:null-pointer)) stats))))
      (values
       (loop for i from 0 below (cffi:foreign-slot-value stats '(:struct bmc-stats) 'stored)
             for p = (cffi:mem-aptr recs '(:struct bmc-record) i)
             collect (list (cffi:foreign-slot-value p '(:struct bmc-record) 'lo)
                           (cffi:foreign-slot-value p '(:struct bmc-record) 'hi)))
       ;; the indices the job used: the next job continues after them (one GPU: a contiguous range;
       ;; several: every device consumed part of its own 2^40-index range, so skip them all)
       (if (rest *bmc-devices*)
           (ash (length *bmc-devices*) 40)
           (cffi:foreign-slot-value stats '(:struct bmc-stats) 'raw))))))

(defun shuffle-list (lst)
  "The kernel stores a wave's deals in no particular order and whole waves at a time: shuffle, so that a
caller who uses fewer deals than a call produced still sees an unbiased sample."
  (let ((v (coerce lst 'vector)))
    (loop for i from (1- (length v)) downto 1
          do (rotatef (aref v i) (aref v (random (1+ i)))))
    (coerce v 'list)))

(defmethod refill ((this deal) n)
  (let ((gen (gpu-gen this)))
    (multiple-value-bind (recs used)
        (sample-records (gpu-gen-handle gen) (gpu-gen-seed gen) (gpu-gen-next-index gen) n)
      (incf (gpu-gen-next-index gen) used)
      (shuffle-list recs))))

(defun deal-result (this hands)
  (with-slots (const-hands defs) this
    (if (and (= (length const-hands) 3) (= (length defs) 1))   ; bridge.cl:1316-1319
        (list (fourth hands))
        hands)))

(defmethod build ((this deal))
  "bridge.cl:1292-1319: one deal as hand objects -- const hands, :~ seats in def order, random hands."
  (let ((gen (gpu-gen this)))
    (unless (gpu-gen-buffer gen) (setf (gpu-gen-buffer gen) (refill this 2048)))
    (deal-result this (apply #'record->hands (pop (gpu-gen-buffer gen))))))

;;; ---- sim (bridge.cl:3144-3162): one batched library call per `sim` -----------------------------------
(defun gpu-dealgen-p (form)
  "The dealgen forms the REST layer and the README use: (build D) or (reorder (build D) 'PERM) with D a
deal object (rest.cl:271, README.md:117-133).  Returns (values deal perm) or nil."
  (cond ((and (consp form) (eq (first form) 'build) (typep (second form) 'deal))
         (values (second form) nil))
        ((and (consp form) (eq (first form) 'reorder)
              (consp (second form)) (eq (first (second form)) 'build) (typep (second (second form)) 'deal))
         (values (second (second form))
                 (let ((p (third form))) (if (and (consp p) (eq (first p) 'quote)) (second p) p))))))

(defmethod sim :around ((self mc-case) runs)
  "Deal generators backed by the library produce all RUNS deals in ONE call (n_accept = runs) instead of
a producer thread, a consumer thread and a (sleep 1) poll loop; the rows (hands m1 m2 ...) are complete
when sim returns.  Any other dealgen form takes the reference path."
  (with-slots (props dealgen trump data) self
    (multiple-value-bind (deal perm) (gpu-dealgen-p dealgen)
      (if (not deal)
          (call-next-method)
          (let* ((recs (subseq (refill deal runs) 0 runs))
                 (rows (loop for (lo hi) in recs
                             for hands = (let ((h (deal-result deal (record->hands lo hi))))
                                           (if perm (reorder h perm) h))
                             collect (cons hands (mapcar (lambda (prop) (apply prop (cons trump hands))) props)))))
            (if data (setf (cdr (last data)) rows) (setf data rows))
            self)))))

;;; ---- tallies -> hist-base (bridge.cl:1325-1330) ----------------------------------------------------------
(defun tallies->base (tallies what seat &optional suit)
  "A bmc_tallies buffer as hist-base's pre-counted BASE argument ((value count) ...): WHAT = :hcp | :len |
:shape (sorted-shape class).  (hist-base more-values (tallies->base ...)) and hist-breakdown /
hist-percent then work unchanged (KAT bridge.cl:1340-1342)."
  (multiple-value-bind (offset bins)
      (ecase what
        (:hcp (values (+ +tally-hcp+ (* 40 seat)) 38))
        (:len (values (+ +tally-len+ (* 14 (+ (* 4 seat) suit))) 14))
        (:shape (values (+ +tally-shape+ (* 40 seat)) 39)))
    (loop for v from 0 below bins
          for n = (cffi:mem-aref tallies :uint64 (+ offset v))
          when (> n 0) collect (list v n))))

(defun sample-tallies (deal-or-handle n-raw &key (seed 1) (first-index 0) (tally-mask 7))
  "Tally-only job: N-RAW indices, no records.  Returns a foreign buffer of +tally-words+ uint64 (free it with
cffi:foreign-free) -- feed it to tallies->base."
  (let ((handle (if (typep deal-or-handle 'deal) (gpu-gen-handle (gpu-gen deal-or-handle)) deal-or-handle))
        (tallies (cffi:foreign-alloc :uint64 :count +tally-words+)))
    (cffi:with-foreign-object (stats '(:struct bmc-stats))
      (with-gpu-call
        (bmc-check (bmc-sample (bmc-ctx) handle seed first-index n-raw 0 0 tally-mask
                               (cffi:null-pointer) 0 tallies stats))))
    tallies))

(defun records-hist-base (records measures)
  "hist-base over measure tuples evaluated on the device (bmc_hist_base): RECORDS = list of (lo hi), MEASURES =
list of (kind seat [suit]) with kind 0 HCP 1 suit length 2 suit HCP 3 losers 4 balanced? 5 line-hcp 6 trump-length
7 sorted-shape class 8 longest-suit.  Returns hist-base's rows (((m1 .. mk) count) ...) in first-seen order."
  (let ((n (length records)) (k (length measures)) (cap 65536))
    (cffi:with-foreign-objects ((recs '(:struct bmc-record) (max n 1)) (ms '(:struct bmc-measure) k)
                                (tuples :uint8 (* cap k)) (counts :uint64 cap) (nu :uint64))
      (loop for (lo hi) in records for i from 0
            for p = (cffi:mem-aptr recs '(:struct bmc-record) i)
            do (setf (cffi:foreign-slot-value p '(:struct bmc-record) 'lo) lo
                     (cffi:foreign-slot-value p '(:struct bmc-record) 'hi) hi))
      (loop for (kind seat suit) in measures for j from 0
            for p = (cffi:mem-aptr ms '(:struct bmc-measure) j)
            do (setf (cffi:foreign-slot-value p '(:struct bmc-measure) 'kind) kind
                     (cffi:foreign-slot-value p '(:struct bmc-measure) 'seat) seat
                     (cffi:foreign-slot-value p '(:struct bmc-measure) 'suit) (or suit 0)
                     (cffi:foreign-slot-value p '(:struct bmc-measure) 'reserved) 0))
      (with-gpu-call (bmc-check (bmc-hist-base (bmc-ctx) recs n ms k tuples counts cap nu)))
      (loop for i from 0 below (cffi:mem-ref nu :uint64)
            collect (list (loop for j from 0 below k collect (cffi:mem-aref tuples :uint8 (+ (* i k) j)))
                          (cffi:mem-aref counts :uint64 i))))))

;;; ---- bidding.cl entry points through bmc_filter -------------------------------------------------------
(defvar *schemes* (make-hash-table :test 'equal :synchronized t)
  "rule table (as printed) -> bmc_scheme handle; tables are few (openings, nt-responses, ...) and live for the session")

(defun gpu-scheme (table)
  (let ((text (let ((*print-pretty* nil)) (prin1-to-string table))))
    (or (gethash text *schemes*)
        (cffi:with-foreign-object (out :pointer)
          (with-gpu-call (bmc-check (bmc-scheme-create text out)))
          (setf (gethash text *schemes*) (cffi:mem-ref out :pointer))))))

(defun filter-hand (hand scheme)
  "bmc_filter on the one-deal record of HAND.  Returns (values lens power losers hcp balanced bid)."
  (multiple-value-bind (lo hi) (hand->record hand)
    (cffi:with-foreign-objects ((rec '(:struct bmc-record)) (feats '(:struct bmc-seat-features) 4))
      (setf (cffi:foreign-slot-value rec '(:struct bmc-record) 'lo) lo
            (cffi:foreign-slot-value rec '(:struct bmc-record) 'hi) hi)
      (with-gpu-call
        (bmc-check (bmc-filter (bmc-ctx) (cffi:null-pointer) (or scheme (cffi:null-pointer)) rec 1
                               (cffi:null-pointer) feats)))
      (let ((f (cffi:mem-aptr feats '(:struct bmc-seat-features) 0)))       ; seat 0 = the hand
        (flet ((bytes (slot) (loop for i from 0 to 3
                                   collect (cffi:mem-aref (cffi:foreign-slot-pointer f '(:struct bmc-seat-features) slot) :uint8 i))))
          (values (bytes 'len) (bytes 'suit-hcp)
                  (cffi:foreign-slot-value f '(:struct bmc-seat-features) 'losers)
                  (cffi:foreign-slot-value f '(:struct bmc-seat-features) 'hcp)
                  (/= 0 (cffi:foreign-slot-value f '(:struct bmc-seat-features) 'balanced))
                  (cffi:foreign-slot-value f '(:struct bmc-seat-features) 'bid)))))))

(defun assess-hand (hand &optional measure)
  "bidding.cl:33-38: (lens power loosers), or (apply MEASURE lens power loosers)."
  (multiple-value-bind (lens power losers) (filter-hand hand nil)
    (let ((vals (list lens power losers)))
      (if measure (apply measure vals) vals))))

(defun bid-index (hand table)
  "Row of TABLE whose rule is the first to hold for HAND (choose-bid's search), nil if none; signals the
reference's error where evaluation would reach an illegal form (bidding.cl:249) before finding one."
  (let ((bid (nth-value 5 (filter-hand hand (gpu-scheme table)))))
    (cond ((= bid -2) (error "illegal function call in a rule of ~A" (mapcar #'car table)))
          ((< bid 0) nil)
          (t bid))))

(defun test-bid (bid hand)
  "bidding.cl:111-112: does BID's rule in `openings` hold for HAND?  (one-row table: only that rule is evaluated)"
  (and (bid-index hand (list (cons bid (open-bid bid)))) t))

(defun choose-bid (hand &optional (openings openings))
  "bidding.cl:146-150: the first bid of the table whose rule holds."
  (let ((i (bid-index hand openings)))
    (and i (car (nth i openings)))))

;;; ---- scoring: contract-score (bridge.cl:3045-3069), base-score, match-points (compscore.cl:94-131) -----------
(defun suit-code (suit)
  (cond ((find suit '(c ♣)) 0) ((find suit '(d ♦)) 1) ((find suit '(h ♥)) 2) ((find suit '(s ♠)) 3) (t 4)))

(defun score-1 (level suit-code declarer-code dbl-code tricks vulnerable)
  (cffi:with-foreign-objects ((c '(:struct bmc-contract)) (v :uint8) (tr :uint8) (out :int32))
    (setf (cffi:foreign-slot-value c '(:struct bmc-contract) 'level) level
          (cffi:foreign-slot-value c '(:struct bmc-contract) 'suit) suit-code
          (cffi:foreign-slot-value c '(:struct bmc-contract) 'declarer) declarer-code
          (cffi:foreign-slot-value c '(:struct bmc-contract) 'dbl) dbl-code
          (cffi:mem-ref v :uint8) (if vulnerable 1 0)
          (cffi:mem-ref tr :uint8) tricks)
    (with-gpu-call (bmc-check (bmc-score (bmc-ctx) c v 1 tr 1 out)))
    (cffi:mem-ref out :int32)))

(defun dbl-code (double)
  (cond ((or (eq double 'xx) (equalp double "XX")) 2) (double 1) (t 0)))

(defun contract-score (suit tricks &key vulnerable double level)
  "bridge.cl:3045-3069, same lambda list; level nil = max(tricks - 6, 1) (0 on the C side)."
  (score-1 (or level 0) (suit-code suit) -1 (dbl-code double) tricks vulnerable))

(defmethod base-score ((this contract) &key vulnerable)
  "compscore.cl:94-100: the contract's score, negated for an E/W declarer (the library applies the sign)."
  (with-slots (level suit declarer double? outcome) this
    (score-1 level (suit-code suit)
             (or (position declarer '(n e s w)) -1)
             (dbl-code double?)
             (+ level 6 (if (eq outcome '=) 0 (if (stringp outcome) (read-from-string outcome) outcome)))
             vulnerable)))

(defmethod match-points ((a contract) (b contract) &key vulnerable)
  "compscore.cl:127-131: IMPs of A against B; off the IMP table (|difference| >= 4000) the reference's
(* nil ...) is a type error -- signalled here too."
  (let ((sa (base-score a :vulnerable (find vulnerable '(ns both))))
        (sb (base-score b :vulnerable (find vulnerable '(ew both)))))
    (cffi:with-foreign-objects ((pa :int32) (pb :int32) (imps :int8) (status :uint8))
      (setf (cffi:mem-ref pa :int32) sa (cffi:mem-ref pb :int32) sb)
      (with-gpu-call (bmc-check (bmc-match-points (bmc-ctx) pa pb 1 imps status)))
      (if (/= 0 (cffi:mem-ref status :uint8))
          (error 'type-error :datum nil :expected-type 'number)
          (cffi:mem-ref imps :int8)))))

;;; ---- compscore tallies over a constrained sample in one pass (BMC_TALLY_SCORE) ------------------------------
(defun sample-score-tallies (deal contracts vulnerable pairs n-accept &key (seed 1))
  "CONTRACTS: contract objects; VULNERABLE: list of booleans; PAIRS: list of (a b) indices.  Samples N-ACCEPT deals
of DEAL and returns (values tricks-hist imps-hist dropped): ((contract (tricks count) ...) ...), ((pair (imps count)
...) ...) ready for hist-breakdown, and the per-pair count of differences off the IMP table."
  (let* ((gen (gpu-gen deal)) (handle (gpu-gen-handle gen)) (nc (length contracts)) (np (length pairs)))
    (cffi:with-foreign-objects ((cs '(:struct bmc-contract) nc) (vul :uint8 nc) (pr :uint8 (max 1 (* 2 np)))
                                (tallies :uint64 +tally-words+) (stats '(:struct bmc-stats)))
      (loop for c in contracts for v in vulnerable for i from 0
            for p = (cffi:mem-aptr cs '(:struct bmc-contract) i)
            do (with-slots (level suit declarer double?) c
                 (setf (cffi:foreign-slot-value p '(:struct bmc-contract) 'level) level
                       (cffi:foreign-slot-value p '(:struct bmc-contract) 'suit) (suit-code suit)
                       (cffi:foreign-slot-value p '(:struct bmc-contract) 'declarer) (or (position declarer '(n e s w)) -1)
                       (cffi:foreign-slot-value p '(:struct bmc-contract) 'dbl) (dbl-code double?)
                       (cffi:mem-aref vul :uint8 i) (if v 1 0))))
      (loop for (a b) in pairs for i from 0
            do (setf (cffi:mem-aref pr :uint8 (* 2 i)) a (cffi:mem-aref pr :uint8 (+ 1 (* 2 i))) b))
      (with-gpu-call
        (bmc-check (bmc-constraints-set-scoring handle cs vul nc pr np))
        (bmc-check (bmc-sample (bmc-ctx) handle seed (gpu-gen-next-index gen) 0 n-accept 16384 8
                               (cffi:null-pointer) 0 tallies stats)))
      (incf (gpu-gen-next-index gen) (cffi:foreign-slot-value stats '(:struct bmc-stats) 'raw))
      (flet ((rows (base stride bins shift)
               (lambda (i)
                 (loop for v from 0 below bins
                       for n = (cffi:mem-aref tallies :uint64 (+ base (* stride i) v))
                       when (> n 0) collect (list (- v shift) n)))))
        (values (loop for i from 0 below nc collect (cons (nth i contracts) (funcall (rows +tally-tricks+ 14 14 0) i)))
                (loop for i from 0 below np collect (cons (nth i pairs) (funcall (rows +tally-imps+ 49 49 24) i)))
                (loop for i from 0 below np collect (cffi:mem-aref tallies :uint64 (+ +tally-imps+ 392 i))))))))

;;; ---- play-deal / best-tricks (bridge.cl:2961-2963, 3027-3028): the trick estimator on the GPU ----------------
(cffi:defcstruct bmc-play-result
  (status :uint8) (n-tricks :uint8) (score :uint8 :count 2) (cards :uint8 :count 52) (winner :uint8 :count 13)
  (reserved :uint8 :count 3))
(cffi:defcfun "bmc_play_hands" :int
  (ctx :pointer) (hands :pointer) (n :uint64) (trump :pointer) (per-deal :int) (arena-kb :uint32) (out :pointer))
(cffi:defcfun "bmc_best_tricks" :int
  (ctx :pointer) (records :pointer) (n :uint64) (trump :pointer) (declarer :pointer) (per-deal :int)
  (arena-kb :uint32) (tricks :pointer))

(defun trump-code (trump)
  "suit symbol (c d h s), suit number, nil or 'nt -> -1 .. 3 (what play-deal's (if (not (eq trump 'nt)) trump) and
normsuit make of it, bridge.cl:1528-1530, 2963)"
  (cond ((null trump) -1) ((numberp trump) trump) (t (or (position trump '(c d h s)) -1))))

(defun play-deal (trump declarer left dummy right)
  "bridge.cl:2961-2963, same lambda list: four hand objects (possibly an ending).  Returns an `outcome` whose tricks,
winner-nos and score slots are what the reference's search returns -- one bmc_play_hands call."
  (cffi:with-foreign-objects ((hands :uint64 4) (tr :int8) (out '(:struct bmc-play-result)))
    (loop for h in (list declarer left dummy right) for k from 0
          do (setf (cffi:mem-aref hands :uint64 k) (hand->lane-mask h)))
    (setf (cffi:mem-ref tr :int8) (trump-code trump))
    (with-gpu-call (bmc-check (bmc-play-hands (bmc-ctx) hands 1 tr 0 0 out)))
    (let ((status (cffi:foreign-slot-value out '(:struct bmc-play-result) 'status))
          (n (cffi:foreign-slot-value out '(:struct bmc-play-result) 'n-tricks))
          (cards (cffi:foreign-slot-pointer out '(:struct bmc-play-result) 'cards))
          (winner (cffi:foreign-slot-pointer out '(:struct bmc-play-result) 'winner))
          (score (cffi:foreign-slot-pointer out '(:struct bmc-play-result) 'score)))
      (case status
        (4 nil)                                                  ; the reference returns nil
        (0 (make-instance 'outcome
             :tricks (loop for tno from 0 below n
                           collect (loop for i from 0 below 4
                                         for c = (cffi:mem-aref cards :uint8 (+ (* 4 tno) i))
                                         collect (if (= c 255) nil (list (ash c -4) (logand c 15)))))
             :winner-nos (loop for tno from 0 below n collect (cffi:mem-aref winner :uint8 tno))
             :winners (loop for tno from 0 below n
                            for w = (cffi:mem-aref winner :uint8 tno)
                            for c = (cffi:mem-aref cards :uint8 (+ (* 4 tno) w))
                            collect (list (ash c -4) (logand c 15)))
             :score (list (cffi:mem-aref score :uint8 0) (cffi:mem-aref score :uint8 1))
             :remaining nil))
        (t (error "play-deal: status ~A (~A)" status (bmc-last-error)))))))

(defun best-tricks (trump n e s w)
  "bridge.cl:3027-3028"
  (second (score (play-deal trump n e s w))))

(defun records-best-tricks (records n trump declarer)
  "best-tricks for a whole record buffer (what an `msr-<contract>` measure over an mc-case needs, rest.cl:287-303):
RECORDS = foreign bmc-record array, TRUMP / DECLARER the same for every deal.  Returns a list of N trick counts."
  (cffi:with-foreign-objects ((tr :int8) (de :uint8) (tricks :uint8 n))
    (setf (cffi:mem-ref tr :int8) (trump-code trump) (cffi:mem-ref de :uint8) declarer)
    (with-gpu-call (bmc-check (bmc-best-tricks (bmc-ctx) records n tr de 0 0 tricks)))
    (loop for i from 0 below n collect (cffi:mem-aref tricks :uint8 i))))

;;; ---- packed index records: about one byte per accepted deal on the way to the host ---------------------------------
(cffi:defcfun "bmc_sample_packed" :int
  (ctx :pointer) (cons :pointer) (seed :uint64) (first-index :uint64) (n-raw :uint64) (n-accept :uint64) (wave :uint64)
  (tally-mask :uint32) (bytes :pointer) (cap-bytes :uint64) (n-bytes :pointer) (tallies :pointer) (stats :pointer))
(cffi:defcfun "bmc_unpack_indices" :int
  (bytes :pointer) (n-bytes :uint64) (offsets :pointer) (cap :uint64) (n :pointer))

(defun sample-packed-offsets (handle seed first-index n-accept &key (wave 262144))
  "N-ACCEPT deals of HANDLE (whole waves of WAVE indices until that many are accepted) as the offsets of their Philox
indices from FIRST-INDEX (a list, chunk order): the deals cross the bus as gap bytes; `bmc_expand` regenerates records
from offsets where they are needed."
  (let* ((cap (+ n-accept wave)) (cap-bytes (+ (* 2 cap) (ash 1 16))))
    (cffi:with-foreign-objects ((bytes :uint8 cap-bytes) (nb :uint64) (offs :uint32 cap) (n :uint64)
                                (tallies :uint64 +tally-words+) (stats '(:struct bmc-stats)))
      (with-gpu-call
        (bmc-check (bmc-sample-packed (bmc-ctx) handle seed first-index 0 n-accept wave 0 bytes cap-bytes nb tallies stats))
        (bmc-check (bmc-unpack-indices bytes (cffi:mem-ref nb :uint64) offs cap n)))
      (loop for i from 0 below (cffi:mem-ref n :uint64) collect (cffi:mem-aref offs :uint32 i)))))
