{-# LANGUAGE BangPatterns #-}
-- | Drop-in for "Data.Text.AhoCorasick.Automaton" (reference: src/Data/Text/AhoCorasick/Automaton.hs).
-- Same export list (:32-44) minus the debug helpers; the state-transition loop (:442-534) runs in
-- libam_b200.so on a B200.  NOT COMPILED HERE (no GHC in the image) -- see INTEGRATION.md.
module Data.Text.AhoCorasick.Automaton
  ( AcMachine (..), CaseSensitivity (..), CodeUnitIndex (..), Match (..), Next (..)
  , build, buildWithCase, runText, runLower, runWithCase
  ) where

import Control.Monad (when)
import Data.Text.CaseSensitivity (CaseSensitivity (..))
import Data.Text.Utf8 (CodeUnitIndex (..), Text (..))
import Foreign.ForeignPtr (ForeignPtr, newForeignPtr, withForeignPtr)
import Foreign.Marshal (alloca, allocaArray, peekArray, withArray)
import Foreign.Ptr (nullPtr)
import Foreign.Storable (peek)
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.Char as Char
import qualified Data.Text.Utf8 as Utf8
import qualified Data.Vector as Vector

import Data.Text.AhoCorasick.FFI

data Match v = Match { matchPos :: {-# UNPACK #-} !CodeUnitIndex, matchValue :: v }   -- Automaton.hs:98-105
data Next a = Done !a | Step !a                                                        -- :398

-- | The device image (immutable, shareable) plus the host-side payload table: the native side reports
-- needle indices, `machineValues` maps them back to the caller's `v`.
data AcMachine v = AcMachine
  { machineHandle :: !(ForeignPtr AmAutomaton)
  , machineCase   :: !CaseSensitivity
  , machineValues :: !(Vector.Vector v)
  }

instance Functor AcMachine where                     -- reference derives Functor (:123)
  fmap f m = m { machineValues = Vector.map f (machineValues m) }

-- | `build :: [(Text, v)] -> AcMachine v` (:176).  The reference builds one machine for both case modes;
-- the device image depends on the mode, so `build` = CaseSensitive and `buildWithCase` picks the mode
-- (Searcher/Replacer below always know it).
build :: [(Text, v)] -> AcMachine v
build = buildWithCase CaseSensitive

buildWithCase :: CaseSensitivity -> [(Text, v)] -> AcMachine v
buildWithCase cs needlesWithValues = unsafePerformIO $
  withArray (map (toSlice . fst) needlesWithValues) $ \needles ->
  withLowerTable cs $ \lowerPtr ->
  alloca $ \out -> do
    rc <- c_am_automaton_build needles (fromIntegral (length needlesWithValues)) (caseToC cs) lowerPtr nullPtr out
    when (rc /= amOk) $ amError "am_automaton_build"
    h <- peek out >>= newForeignPtr c_am_automaton_free_ptr
    pure $ AcMachine h cs (Vector.fromList (map snd needlesWithValues))

-- | `runWithCase` (:443): left fold over all matches in the reference's order, honouring `Done`.
runWithCase :: CaseSensitivity -> a -> (a -> Match v -> Next a) -> AcMachine v -> Text -> a
runWithCase cs seed f machine text
  | cs /= machineCase machine = error "AcMachine was built for the other CaseSensitivity"
  | otherwise = go seed (findAll machine text)
  where
    go !acc [] = acc
    go !acc (AmMatch pos i : ms) =
      case f acc (Match (CodeUnitIndex (fromIntegral pos)) (machineValues machine `Vector.unsafeIndex` fromIntegral i)) of
        Step acc' -> go acc' ms
        Done r    -> r

runText, runLower :: a -> (a -> Match v -> Next a) -> AcMachine v -> Text -> a
runText  = runWithCase CaseSensitive   -- :539-541
runLower = runWithCase IgnoreCase      -- :551-553

-- | am_find_all with the overflow protocol: retry once with the exact size.
findAll :: AcMachine v -> Text -> [AmMatch]
findAll machine text = unsafePerformIO $
  withForeignPtr (machineHandle machine) $ \h -> withSlice text $ \hay -> alloca $ \nFound ->
    let attempt cap = allocaArray cap $ \buf -> do
          rc <- c_am_find_all h hay buf (fromIntegral cap) nFound
          n <- fromIntegral <$> peek nFound
          if rc == amEOverflow then attempt n
          else if rc /= amOk then amError "am_find_all"
          else peekArray n buf
    in attempt 4096
  -- withSlice / toSlice pin the array if needed (Utf8.isArrayPinned, :325-331) and build a U8Slice;
  -- withLowerTable enumerates [(c, Char.toLower c) | c <- [chr 128 ..], Char.toLower c /= c] once (a CAF).
