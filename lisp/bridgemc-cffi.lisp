;;; bridgemc-cffi.lisp -- CFFI shim that puts libbridgemc.so behind bridge-mc's
;;; Lisp entry points (see INTEGRATION.md).  Load AFTER bridge.cl / bidding.cl /
;;; compscore.cl.  NOT EXECUTED in the build image (no Lisp implementation there);
;;; the same C ABI is exercised call for call by bridge_mc_b200/ (ctypes).

(ql:quickload :cffi)

(cffi:define-foreign-library libbridgemc (t (:default "libbridgemc")))
(cffi:use-foreign-library libbridgemc)

(cffi:defcstruct bmc-record (lo :uint64) (hi :uint64))
(cffi:defcstruct bmc-seat-spec
  (hcp :int8 :count 2) (len :int8 :count 8) (suit-hcp :int8 :count 8) (losers :int8 :count 2)
  (fixed :uint8) (reserved :uint8 :count 3) (fixed-mask :uint64))
(cffi:defcstruct bmc-contract (level :int8) (suit :int8) (declarer :int8) (dbl :int8))
(cffi:defcstruct bmc-stats
  (raw :uint64) (voided :uint64) (accepted :uint64) (stored :uint64)
  (waves :uint32) (launches :uint32) (kernel-ms :float) (total-ms :float))

(cffi:defcfun "bmc_init" :int (device :int) (out :pointer))
(cffi:defcfun "bmc_shutdown" :void (ctx :pointer))
(cffi:defcfun "bmc_last_error" :string)
(cffi:defcfun "bmc_constraints_create" :int (specs :pointer) (rules :pointer) (out :pointer))
(cffi:defcfun "bmc_constraints_destroy" :void (c :pointer))
(cffi:defcfun "bmc_constraints_set_reference_order" :int (c :pointer) (n-defs :int))
(cffi:defcfun "bmc_constraints_set_mode" :int (c :pointer) (mode :int))
(cffi:defcfun "bmc_constraints_set_jit" :int (c :pointer) (on :int))
;; what determines the deal behind a Philox index besides the seed (log them with a reproducible job)
(cffi:defcfun "bmc_constraints_mode" :int (c :pointer))
(cffi:defcfun "bmc_constraints_primary" :int (c :pointer))
(cffi:defcfun "bmc_constraints_primary_hcp" :int (c :pointer) (lo :pointer) (hi :pointer))
(cffi:defcfun "bmc_hist_base" :int
  (ctx :pointer) (records :pointer) (n :uint64) (measures :pointer) (k :uint32)
  (tuples :pointer) (counts :pointer) (cap :uint64) (n-unique :pointer))
(cffi:defcfun "bmc_scheme_create" :int (table :string) (out :pointer))
(cffi:defcfun "bmc_scheme_destroy" :void (s :pointer))
(cffi:defcfun "bmc_sample" :int
  (ctx :pointer) (cons :pointer) (seed :uint64) (first-index :uint64) (n-raw :uint64)
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

(defvar *bmc-ctx* nil)
(defun bmc-ctx ()
  (or *bmc-ctx*
      (cffi:with-foreign-object (out :pointer)
        (bmc-check (bmc-init 0 out))
        (setf *bmc-ctx* (cffi:mem-ref out :pointer)))))

(defun bmc-check (rc)
  (cond ((= rc 0) t)
        ((= rc -3) (error 'bad-range :msg (bmc-last-error) :down nil :up nil))   ; bridge.cl:1109-1123
        (t (error "libbridgemc: ~A (~A)" (bmc-last-error) rc))))

;; --- records -> hand objects (bridge.cl:1622-1633) --------------------------
(defun record->hands (lo hi)
  (loop for seat from 0 to 3
        collect (mkhand
                 (loop for suit from 0 to 3
                       collect (loop for rank from 0 to 12
                                     for bit = (+ (* 16 suit) rank)
                                     when (and (eq (logbitp bit lo) (logbitp 0 seat))
                                               (eq (logbitp bit hi) (logbitp 1 seat)))
                                       collect rank)))))

;; --- range defs -> bmc_seat_spec --------------------------------------------
(defun fill-range (ptr slot index spec)
  "spec: number | (lo hi ...) | nil ; -1 = nil"
  (let ((lo (cond ((numberp spec) spec) ((listp spec) (first spec))))
        (hi (cond ((numberp spec) spec) ((listp spec) (second spec)))))
    (setf (cffi:mem-aref (cffi:foreign-slot-pointer ptr '(:struct bmc-seat-spec) slot) :int8 (* 2 index)) (or lo -1)
          (cffi:mem-aref (cffi:foreign-slot-pointer ptr '(:struct bmc-seat-spec) slot) :int8 (+ 1 (* 2 index))) (or hi -1))))

(defun hand->lane-mask (hand)
  (loop for suit from 0 to 3
        sum (loop for rank in (nth suit (suits hand)) sum (ash 1 (+ (* 16 suit) rank)))))

;; --- class gpu-deal: `build` keeps its signature (bridge.cl:1292) -------------
(defclass gpu-deal (deal)
  ((handle :initform nil) (buffer :initform nil) (next-index :initform 0)
   (seed :initform (random (ash 1 62)))))

(defmethod gpu-handle ((this gpu-deal))
  (with-slots (handle defs const-hands) this
    (or handle
        (cffi:with-foreign-objects ((specs '(:struct bmc-seat-spec) 4) (out :pointer))
          (loop for k from 0 to 3
                for p = (cffi:mem-aptr specs '(:struct bmc-seat-spec) k)
                do (loop for i from 0 below (cffi:foreign-type-size '(:struct bmc-seat-spec))
                         do (setf (cffi:mem-aref p :uint8 i) 0))
                   (fill-range p 'hcp 0 nil) (fill-range p 'losers 0 nil)
                   (dotimes (s 4) (fill-range p 'len s nil) (fill-range p 'suit-hcp s nil)))
          (loop for hand in const-hands
                for k from 0
                for p = (cffi:mem-aptr specs '(:struct bmc-seat-spec) k)
                do (setf (cffi:foreign-slot-value p '(:struct bmc-seat-spec) 'fixed) 1
                         (cffi:foreign-slot-value p '(:struct bmc-seat-spec) 'fixed-mask)
                         (hand->lane-mask (mkhand hand))))
          (loop for def in defs
                for k from (length const-hands)
                for p = (cffi:mem-aptr specs '(:struct bmc-seat-spec) k)
                do (destructuring-bind (&key hcp c d h s &allow-other-keys) def
                     (fill-range p 'hcp 0 hcp)
                     (loop for spec in (list c d h s) for i from 0
                           do (fill-range p 'len i spec)
                              (fill-range p 'suit-hcp i (and (listp spec) (third spec))))))
          (bmc-check (bmc-constraints-create specs (cffi:null-pointer) out))
          (setf handle (cffi:mem-ref out :pointer))
          ;; keep `build`'s own seat-by-seat distribution (and quirks); no def still samples one seat
          (bmc-check (bmc-constraints-set-reference-order handle (max 1 (length defs))))
          handle))))

(defmethod refill ((this gpu-deal) n)
  (with-slots (next-index seed) this
    (let ((cap (+ n 16384)))
      (cffi:with-foreign-objects ((recs '(:struct bmc-record) cap) (stats '(:struct bmc-stats)))
        (sb-int:with-float-traps-masked (:invalid :divide-by-zero :overflow)
          (bmc-check (bmc-sample (bmc-ctx) (gpu-handle this) seed next-index 0 n 16384 0
                                 recs cap (cffi:null-pointer) stats)))
        (incf next-index (* 16384 (cffi:foreign-slot-value stats '(:struct bmc-stats) 'waves)))
        (loop for i from 0 below (cffi:foreign-slot-value stats '(:struct bmc-stats) 'stored)
              for p = (cffi:mem-aptr recs '(:struct bmc-record) i)
              collect (list (cffi:foreign-slot-value p '(:struct bmc-record) 'lo)
                            (cffi:foreign-slot-value p '(:struct bmc-record) 'hi)))))))

(defmethod build ((this gpu-deal))
  (with-slots (buffer const-hands defs) this
    (unless buffer (setf buffer (refill this 2048)))
    (let ((hands (apply #'record->hands (pop buffer))))
      (if (and (= (length const-hands) 3) (= (length defs) 1))   ; bridge.cl:1316-1319
          (list (fourth hands))
          hands))))

;; --- scoring pass-throughs (bridge.cl:3045, compscore.cl:94) -------------------
(defun suit-code (suit)
  (cond ((find suit '(c ♣)) 0) ((find suit '(d ♦)) 1) ((find suit '(h ♥)) 2) ((find suit '(s ♠)) 3) (t 4)))

(defun gpu-contract-score (suit tricks &key vulnerable double level)
  (cffi:with-foreign-objects ((c '(:struct bmc-contract)) (v :uint8) (tr :uint8) (out :int32))
    (setf (cffi:foreign-slot-value c '(:struct bmc-contract) 'level) (or level 0)
          (cffi:foreign-slot-value c '(:struct bmc-contract) 'suit) (suit-code suit)
          (cffi:foreign-slot-value c '(:struct bmc-contract) 'declarer) -1
          (cffi:foreign-slot-value c '(:struct bmc-contract) 'dbl) (cond ((eq double 'xx) 2) (double 1) (t 0))
          (cffi:mem-ref v :uint8) (if vulnerable 1 0)
          (cffi:mem-ref tr :uint8) tricks)
    (bmc-check (bmc-score (bmc-ctx) c v 1 tr 1 out))
    (cffi:mem-ref out :int32)))

;; --- rule tables: openings, (nt-responses n), ... go over as printed text -------
(defun gpu-scheme (table)
  (cffi:with-foreign-object (out :pointer)
    (bmc-check (bmc-scheme-create (let ((*print-pretty* nil)) (prin1-to-string table)) out))
    (cffi:mem-ref out :pointer)))
