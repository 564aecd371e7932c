/* Minimal C caller of the host-buffer front end (include/recnow_b200.h): one batch in host memory in, loss / pair count /
 * gradient in host memory out.  Plain C99; link against rec_now_b200/librecnow_b200.so:
 *   gcc -std=c99 -Iinclude examples/host_pairwise.c -o host_pairwise -Lrec_now_b200 -lrecnow_b200 -Wl,-rpath,$PWD/rec_now_b200
 * Replaces a call of pairwise_loss(outputs, labels, groups, click_occurance_power=-0.5)
 * (rec_now/rec_block/pairwise_loss_from_batch.py:228-279) + the gradient TF autodiff would return. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "recnow_b200.h"

int main(void) {
  enum { B = 5 };
  /* the batch of tests/rec_block/test_pairwise_loss_from_batch.py:33-48 */
  int64_t keys[B] = {1, 1, 2, 2, 2};
  float logits[B] = {0.f, 1.f, 2.f, 3.f, 4.f}, labels[B] = {1.1f, 0.f, 0.f, 1.f, 1.f};
  float loss = 0.f, n_f32 = 0.f, dlogits[B];
  int64_t n_pair = 0;
  printf("librecnow_b200 version %d\n", rn_version());

  rn_host_pairwise* hp = NULL;
  int rc = rn_host_pairwise_create(B, 1, 2, &hp);
  if (rc != RN_OK) { printf("create: %s (needs a B200)\n", rn_strerror(rc)); return rc == RN_ERR_NO_DEVICE || rc == RN_ERR_LAUNCH ? 0 : 1; }

  rn_pairwise_args a;
  memset(&a, 0, sizeof a);
  a.B = B; a.K = 1; a.label_func = RN_LABEL_STEP;
  a.keys = keys; a.logits = logits; a.labels = labels;
  a.factor = 1.0f; a.power = -0.5f; a.reduce_mean = 1; a.part_rank = 0; a.part_count = 1;
  a.loss = &loss; a.n_pair_f32 = &n_f32; a.n_pair = &n_pair; a.dlogits = dlogits;

  int32_t ticket = -1;
  rc = rn_host_pairwise_submit(hp, &a, &ticket);
  if (rc == RN_OK) rc = rn_host_pairwise_wait(hp, ticket);
  if (rc != RN_OK) { printf("submit / wait: %s\n", rn_strerror(rc)); rn_host_pairwise_destroy(hp); return 1; }
  printf("n_pair = %lld (reference: 3)  loss = %.7f (reference: 0.5415076)\n", (long long)n_pair, loss);
  for (int i = 0; i < B; ++i) printf("  dloss/dlogit[%d] = % .7f\n", i, dlogits[i]);
  rn_host_pairwise_destroy(hp);
  return 0;
}
