;;; -*- Mode:Lisp; Syntax:ANSI-Common-Lisp; Coding:utf-8 -*-
;;;
;;; llacuda-glue.lisp -- the Lisp side of the llacuda drop-in for tpapp/lla.
;;;
;;; WRITTEN BLIND: the build image has no Common Lisp implementation (sbcl/ecl/ccl/clisp are all
;;; absent) and none of LLA's Lisp dependencies, so nothing in this file has been compiled or run.
;;; It is kept tiny and mechanical on purpose.  What IS tested is the C ABI these forms bind to
;;; (tests/ drive the very same symbols through ctypes with LLA's argument lists, see
;;; lla_b200/lla.py, which mirrors LLA's layers L7..L3 one to one).
;;;
;;; Load AFTER the stock LLA system:   (asdf:load-system "lla") (load "llacuda-glue.lisp")
;;;
;;; Option A of INTEGRATION.md needs none of this: put "libllacuda.so" first in
;;;   (defvar cl-user::*lla-configuration* '(:libraries ("libllacuda.so" "liblapack.so" "libblas.so")))
;;; and the unchanged `foreign-funcall "dgemm_"` of src/fortran-call.lisp:428-439 resolves into it.
;;; This file is Option B: explicit re-pointing of the name generator plus the device-resident
;;; extras, touching the subsystems BASELINE.json names.

(in-package #:lla)

;;;; --- libraries.lisp: load the device library next to the host BLAS/LAPACK -------------------
;;; reference: src/libraries.lisp:10-12 loads (query-configuration :libraries ...) at load time.

(defparameter *llacuda-library* (query-configuration :llacuda-library "libllacuda.so"))
(defparameter *llacuda-enabled* nil
  "T once libllacuda.so is loaded and llacuda_init() found an sm_100 device.")
(defparameter *llacuda-min-dimension* (query-configuration :llacuda-min-n 256)
  "Below this matrix order the PCIe round trip costs more than host BLAS; calls stay on the host.")

(defparameter *llacuda-gpus* (query-configuration :llacuda-gpus 1)
  "Number of GPUs one call is partitioned over (the multi-GPU engine inside libllacuda.so: one process, one host
thread per GPU, NCCL panel broadcasts).  > 1 routes the hot-path routines of order >= 8192 to llacuda_pd<name>.")

(defun llacuda-load ()
  (cffi:load-foreign-library *llacuda-library*)
  (setf *llacuda-enabled*
        (zerop (cffi:foreign-funcall "llacuda_init" :int)))
  (unless *llacuda-enabled*
    (warn "llacuda: ~A" (cffi:foreign-funcall "llacuda_last_error" :string)))
  (when *llacuda-enabled*
    ;; BLAS-style entry points: record the failure and return (LLACUDA-CHECK-BLAS signals it) instead of the
    ;; XERBLA-style stop of the process
    (cffi:foreign-funcall "llacuda_set_error_mode" :int 1 :void)
    (when (> *llacuda-gpus* 1)
      (unless (zerop (cffi:foreign-funcall "llacuda_dist_init_all" :int *llacuda-gpus* :int 0 :int 0 :int))
        (warn "llacuda: multi-GPU engine unavailable: ~A" (cffi:foreign-funcall "llacuda_last_error" :string))
        (setf *llacuda-gpus* 1))))
  *llacuda-enabled*)

(defun llacuda-check-blas ()
  "Call after every BLAS-CALL that went to the device library: the BLAS symbols have no INFO argument, a device-side
failure poisons the output with NaN and is reported here (include/llacuda.h, llacuda_set_error_mode)."
  (unless (zerop (cffi:foreign-funcall "llacuda_last_error_code" :int))
    (let ((message (cffi:foreign-funcall "llacuda_last_error" :string)))
      (cffi:foreign-funcall "llacuda_clear_error" :void)
      (error 'llacuda-device-error :info -1 :message message))))

(defun llacuda-worthwhile-p (&rest dimensions)
  "The size threshold of SURVEY.md section 5 (:cuda-min-n): below *LLACUDA-MIN-DIMENSION* the PCIe round trip costs
more than the host BLAS call, so the name generator keeps the host symbol."
  (and *llacuda-enabled* (every (lambda (d) (>= d *llacuda-min-dimension*)) dimensions)))

;;;; --- fortran-call.lisp: re-point the name generator for the hot-path routines ----------------
;;; reference: src/fortran-call.lisp:405-416 builds "dgemm_" from type letter + name; the ECASE of
;;; :428-439 is generated at macroexpansion time, so LLA's linear-algebra.lisp and blas.lisp must be
;;; recompiled after this redefinition (asdf:load-system "lla" :force t), or use Option A.

(defparameter *llacuda-routines*
  '(("D" . ("gemm" "getrf" "getrs" "gesv" "potrf" "potrs" "syrk" "posv" "trsm" "getri" "potri" "trtri"))
    ("S" . ("gemm" "getrf" "getrs" "gesv" "potrf" "potrs" "syrk" "posv"))
    ("C" . ("gemm" "getrf" "getrs" "gesv" "potrf" "potrs" "herk" "posv"))
    ("Z" . ("gemm" "getrf" "getrs" "gesv" "potrf" "potrs" "herk" "posv")))
  "Routines libllacuda.so exports as llacuda_<letter><name> with the Fortran signature.")

(defun llacuda-routine-p (letter name)
  (member name (cdr (assoc letter *llacuda-routines* :test #'string-equal))
          :test #'string-equal))

(defun blas-lapack-function-name (type name)
  "As the reference (src/fortran-call.lisp:405-416), but emits the llacuda_ entry point for the
routines the device library implements.  Same argument list, same void return, same INFO cell."
  (let+ (((real-name &optional (complex-name name)) (ensure-list name))
         (letter (switch (type)
                   (+single+ "S")
                   (+double+ "D")
                   (+complex-single+ "C")
                   (+complex-double+ "Z")))
         (name (if (complex? type) complex-name real-name)))
    (if (llacuda-routine-p letter name)
        (format nil "llacuda_~(~A~A~)" letter name)
        (format nil "~(~A~A_~)" letter name))))

;;;; --- fortran-call.lisp: the size threshold and the multi-GPU entry points, decided at run time --------
;;; BLAS-LAPACK-FUNCTION-NAME runs at macroexpansion time and cannot see the dimensions of a call, so the threshold
;;; lives in the call form: the generated ECASE branch becomes a run-time choice between the device symbol, the
;;; multi-GPU symbol (llacuda_pd<name>: same argument list) and the host symbol.  *LLACUDA-CALL-DIMENSION* is bound by
;;; the hot-path methods (mm, lu, solve, cholesky) around their BLAS-CALL / LAPACK-CALL to the smallest matrix
;;; dimension of the call; unbound (any other caller) means "host".  The library itself has no CPU fallback: the
;;; fallback below is the reference's own host LAPACK, loaded next to libllacuda.so.  (Written blind, see the header.)

(defvar *llacuda-call-dimension* nil)

(defparameter *llacuda-multi-gpu-routines* '("gemm" "getrf" "gesv" "potrf" "potrs" "posv")
  "D (and Z for gemm) routines that libllacuda.so also exports as llacuda_pd<name> / llacuda_pz<name>.")

(defun llacuda-names (type name)
  "Device, multi-GPU (or NIL) and host symbol for TYPE / NAME."
  (let+ (((real-name &optional (complex-name name)) (ensure-list name))
         (letter (switch (type) (+single+ "s") (+double+ "d") (+complex-single+ "c") (+complex-double+ "z")))
         (name (if (complex? type) complex-name real-name))
         (host (format nil "~A~A_" letter name)))
    (values (when (llacuda-routine-p letter name) (format nil "llacuda_~A~A" letter name))
            (when (and (member name *llacuda-multi-gpu-routines* :test #'string-equal)
                       (or (string= letter "d") (and (string= letter "z") (string-equal name "gemm"))))
              (format nil "llacuda_p~A~A" letter name))
            host)))

(defun blas-lapack-call-form (type-var name arguments &optional (return-types :void))
  "As the reference (src/fortran-call.lisp:428-439), with a run-time choice of the library per branch."
  (let ((arguments (arguments-for-cffi arguments)))
    `(ecase ,type-var
       ,@(loop for type in +float-types+
               for return-type in (blas-return-types return-types)
               collect
               (multiple-value-bind (device multi host) (llacuda-names type name)
                 `(,type
                   (cond
                     ,@(when multi
                         `(((and *llacuda-call-dimension* (> *llacuda-gpus* 1) (>= *llacuda-call-dimension* 8192))
                            (cffi:foreign-funcall ,multi ,@arguments ,return-type))))
                     ,@(when device
                         `(((and *llacuda-call-dimension* (llacuda-worthwhile-p *llacuda-call-dimension*))
                            (prog1 (cffi:foreign-funcall ,device ,@arguments ,return-type)
                              (llacuda-check-blas)))))
                     (t (cffi:foreign-funcall ,host ,@arguments ,return-type)))))))))

;;;; --- conditions: device failures are NOT LAPACK INFO codes -----------------------------------
;;; reference: src/fortran-call.lisp:336-347 maps INFO<0 to lapack-invalid-argument.  libllacuda
;;; reports CUDA failures as INFO <= -10000 plus a message; surface them as their own condition.

(define-condition llacuda-device-error (error)
  ((info :initarg :info :reader info)
   (message :initarg :message :reader message))
  (:report (lambda (c stream)
             (format stream "llacuda device failure (INFO=~A): ~A" (info c) (message c)))))

(defun llacuda-check-info (info)
  (when (<= info -10000)
    (error 'llacuda-device-error
           :info info
           :message (cffi:foreign-funcall "llacuda_last_error" :string))))

;;;; --- foreign-memory.lisp / pinned-array.lisp: pinned staging for large operands ---------------
;;; reference: src/pinned-array.lisp:105-109 (WITH-WORK-AREA = cffi:with-foreign-pointer) and
;;; :135-141 (sb-sys:with-pinned-objects).  Lisp heap vectors are pageable; for operands above a
;;; few MiB a page-locked work area lets the library DMA at PCIe rate.  The library detects
;;; page-locked pointers itself, so this is purely an allocation policy.

(defmacro with-llacuda-work-area ((pointer size) &body body)
  "Like WITH-WORK-AREA but the SIZE bytes are page-locked host memory (llacuda_host_alloc)."
  (with-unique-names (holder)
    `(cffi:with-foreign-object (,holder :pointer)
       (let ((rc (cffi:foreign-funcall "llacuda_host_alloc" :pointer ,holder :size ,size :int)))
         (unless (zerop rc) (llacuda-check-info rc)))
       (let ((,pointer (cffi:mem-ref ,holder :pointer)))
         (unwind-protect (progn ,@body)
           (cffi:foreign-funcall "llacuda_host_free" :pointer ,pointer :int))))))

;;;; --- factorizations.lisp / linear-algebra.lisp: device-resident factors (SURVEY 8f rank 1) ----
;;; A LU or CHOLESKY object may keep its factors on the device so that (solve factorization b)
;;; does not re-upload (and, for LU, re-transpose: src/linear-algebra.lisp:264-276) n^2 numbers on
;;; every call.  The handle is a device pointer from llacuda_malloc; the *_dev entry points of
;;; include/llacuda.h take it directly (column-major, 64-bit dimensions).

(defclass device-lu (lu)
  ((device-lu :initarg :device-lu :reader device-lu :documentation "device pointer, column-major LU")
   (device-ipiv :initarg :device-ipiv :reader device-ipiv))
  (:documentation "LU factorization whose factors also live in HBM."))

(defun llacuda-device-free (pointer)
  (cffi:foreign-funcall "llacuda_free" :pointer pointer :int))

;;; (solve device-lu b) => llacuda_dgetrs_dev('N', n, nrhs, dev-lu, n, dev-ipiv, dev-b, n), with b
;;; staged through llacuda_memcpy_h2d / llacuda_dtranspose_dev instead of TRANSPOSE-MATRIX-TO-MEMORY
;;; (src/foreign-memory.lisp:211-230).  The method bodies follow the host methods line by line and
;;; are left to the maintainer who can run them; the C side is complete and tested.

;;;; --- linear-algebra.lisp: row-major-native entry points (SURVEY 8b Option B) -----------------------------
;;; The reference transposes A (and B, X) element by element in Lisp on both sides of the call
;;; (src/foreign-memory.lisp:211-255).  llacuda_dgesv_rm / llacuda_dgetrf_rm / llacuda_dgetrs_rm /
;;; llacuda_dpotrs_rm take the row-major buffers as they are -- every argument a pointer, INFO by pointer, exactly
;;; like the Fortran symbols -- so the method is the reference's LAPACK-CALL form minus the :TRANSPOSE? flags.
;;; Sketch for (solve array array), reference src/linear-algebra.lisp:278-289 (again written blind):
;;;
;;;   (defmethod solve ((a array) (b array))
;;;     (let+ (((a0 a1) (array-dimensions a))
;;;            ((&values b0 b1 &ign) (dimensions-as-matrix b :column)))
;;;       (assert (= a0 a1 b0) () 'lla-incompatible-dimensions)
;;;       (if (and *llacuda-enabled* (eq (common-float-type a b) +double+) (>= a0 *llacuda-min-dimension*))
;;;           (let ((x (copy-array b)))              ; B is input, X a fresh array: one plain copy, no transpose
;;;             (with-pinned-array (pa a) (with-pinned-array (px x)
;;;               (with-fortran-atoms ((n a0) (nrhs b1) (lda a1) (ldb b1) (info 0))
;;;                 (cffi:foreign-funcall "llacuda_dgesv_rm" :pointer n :pointer nrhs :pointer pa :pointer lda
;;;                                       :pointer px :pointer ldb :pointer info :void)
;;;                 (llacuda-check-info (mem-ref info :int32))   ; then the reference's INFO -> condition mapping
;;;                 x))))
;;;           (call-next-method))))
;;;
;;; Measured through the ctypes mirror of this binding (lla_b200/lla.py, native=True) on B200, n = 8192, 64 right-hand
;;; sides: 97.5 ms against 1412 ms for the transposing path, identical X (tools/resident_probe.py).
