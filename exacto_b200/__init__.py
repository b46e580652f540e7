"""exacto_b200 -- B200 (sm_100a) implementation of exacto's ciphertext-multiplication hot path.

Host mirror of the reference's public surface for that path (same names, argument meaning
and error behaviour) over the C ABI of ``libexacto_b200.so`` (include/exacto_b200.h):

    ring/      CoeffPoly, NttPoly, RnsPoly, RnsBasis, make_plan; modular (scalar helpers of ring/modular.rs)
    params/    BfvParamsBuilder, BfvParams, DbfvParams, compact_bfv, compact_dbfv, u64_dbfv
    bfv/       BfvCiphertext, RelinKey, bfv_mul_and_relin, bfv_add, bfv_sub, bfv_neg,
               GaloisKey, bfv_apply_automorphism, bfv_trace, bfv_inner_product
    dbfv/      DbfvCiphertext, dbfv_mul, dbfv_add, dbfv_sub, dbfv_neg, dbfv_apply_automorphism
    encrypt/   SecretKey, encode_scalar, encrypt_sk_with_samples, decrypt, dbfv_encrypt_*_with_samples,
               dbfv_decrypt, dbfv_decrypt_poly (sampling stays with the caller)
    keygen/    gen_secret_key / gen_relin_key / gen_galois_key _with_sampler, apply_automorphism
    bootstrap/ dbfv_mul_then_bootstrap, dbfv_mul_chain_then_bootstrap, dbfv_bootstrap, bfv_bootstrap,
               coeffs_to_slots, slots_to_coeffs, eval_poly_homomorphic (Paterson-Stockmeyer), rounding polynomial

There is no CPU fallback: every operation needs the CUDA library and a GPU.
"""
from .error import ExactoError
from .params import (BfvParams, BfvParamsBuilder, DbfvParams, RnsBasis, cfg3_prime_dbfv, compact_bfv,
                     compact_dbfv, compute_gadget_digits, set_default_device, small_bfv, u64_dbfv)
from . import modular
from .ring import CoeffPoly, NttPoly, Plan, RnsPoly, make_plan
from .bfv import (BfvCiphertext, GaloisKey, RelinKey, bfv_add, bfv_apply_automorphism, bfv_apply_automorphism_batch,
                  bfv_inner_product, bfv_mul_and_relin, bfv_mul_and_relin_batch, bfv_mul_no_relin, bfv_mul_no_relin_batch,
                  gadget_decompose, relinearize, relinearize_batch, bfv_neg, bfv_plain_add, bfv_plain_mul,
                  bfv_scalar_mul, bfv_sub, bfv_trace, scale_plaintext, trivial_encrypt)
from .dbfv import (DbfvCiphertext, dbfv_add, dbfv_apply_automorphism, dbfv_relinearize, dbfv_mul, dbfv_mul_batch, dbfv_neg, dbfv_sub, small_reps)
from .encrypt import (SecretKey, decode_scalar, decrypt, decrypt_batch, dbfv_decrypt, dbfv_decrypt_poly,
                      dbfv_encrypt_poly_sk_with_samples, dbfv_encrypt_sk_with_samples, digit_decompose,
                      digit_recompose_signed, encode_scalar, encrypt_sk_with_samples, encrypt_sk_with_sampler,
                      encrypt_pk_with_sampler, dbfv_encrypt_sk_with_sampler, dbfv_encrypt_with_sampler)
from .keygen import (PublicKey, apply_automorphism, gen_galois_key_with_sampler, gen_public_key_with_sampler,
                     gen_relin_key_with_sampler, gen_secret_key_with_sampler)
from .advanced import dbfv_change_base, dbfv_div_by_base
from .hostmem import PendingMul, PinnedArray, dbfv_mul_batch_async, pinned_empty
from .bootstrap import (BootstrapKey, bfv_bootstrap, bfv_monomial_mul, coeffs_to_slots, compute_rounding_poly,
                        create_boot_sk, dbfv_bootstrap, dbfv_mul_chain_then_bootstrap, dbfv_mul_then_bootstrap,
                        eval_poly_homomorphic, eval_poly_homomorphic_batch, extract_coefficient,
                        gen_bootstrap_key_with_sampler, lagrange_interpolate, required_trace_elements, slots_to_coeffs,
                        trivial_encrypt_poly)

__all__ = [n for n in dir() if not n.startswith("_")]
