import numpy as np, torch, sys
sys.path.insert(0, '.')
import crbm_b200
np.random.seed(0)
K, M, B, L = 100, 15, 100, 200
m = crbm_b200.CRBM(K, M, epochs=1, batchsize=B, cd_k=1, seed=1234)
data = crbm_b200.PackedSequences.random(20000, L, seed=1234)
for step in range(0, 200):
    m._train_packed(data, (step % 200) * B, (step % 200 + 1) * B)
    if step in (0, 1, 2, 3, 5, 8, 12, 16, 24, 32, 48, 64, 100, 150, 199):
        h = m.fantasy_h.get_value(); hp = m.fantasy_h_prime.get_value()
        st = m._train_state['stats'].cpu().numpy(); lay = m._train_state['layout']
        pk = st[lay.h_fwd:lay.h_fwd + K] / st[lay.n_hidden_data]
        print(step, 'fantasy density fwd %.4f rc %.4f' % (h.mean(), hp.mean()), 'data mean p %.4f' % pk.mean(),
              'b mean %.2f' % m.bias.get_value().mean(), 'Wrms %.3f' % np.sqrt((m.motifs.get_value() ** 2).mean()))
